// nn_kernels.cu — leaf evaluation on B200: on-device observation encoding fused with the MLP
// forward pass on 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM, weights streamed by
// the TMA bulk-copy engine into 128B-swizzled shared memory).
//
// Replaces, for the MLP architecture, the reference's leaf evaluators
//   FlatEncoder::encode_into      crates/alpharat-sampling/src/flat_encoder.rs:52-124
//   OnnxBackend::evaluate_batch   crates/alpharat-sampling/src/backends/onnx.rs:176-245
//   TensorrtBackend::evaluate_batch  .../backends/tensorrt.rs:423-535
// whose graph is PyRatMLP.predict (alpharat/nn/models/mlp.py:120-153), eval mode:
//   x[349] -> Linear(349,256) -> BN -> ReLU -> Linear(256,256) -> BN -> ReLU
//          -> {Linear(256,5) -> softmax} x2 ; Linear(256,2) -> softplus
//
// Numerics: observations, weights and hidden activations are bf16, accumulation is fp32
// (tcgen05 kind::f16), heads/softmax/softplus in fp32.  BatchNorm (running stats) is folded
// into the preceding Linear on the host.  Parity vs torch fp32 is a stated tolerance
// (tests/test_gpu_nn.py); parity vs a bf16-emulating torch reference is ~1e-3.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "nn_api.cuh"

namespace ar {

// ---------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> f32, issued by one thread for the CTA
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// 16 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float v[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, 128B-swizzled operand tile: rows of 64 bf16 (128 B), 8-row groups of 1024 B.
// Descriptor per cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO [16,30) = 1, SBO [32,46) = 64,
// version [46,48) = 1, layout [61,64) = 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M x N
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// byte offset of element (row, col) inside a [rows x 64] SW128 tile
__host__ __device__ inline uint32_t sw128_offset(int row, int col) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((col >> 3) ^ (row & 7)) & 7) << 4) + (col & 7) * 2);
}

// ---------------------------------------------------------------------------------------
// Observation encoding (flat_encoder.rs:52-124)
// ---------------------------------------------------------------------------------------
struct RowView {
  uint64_t cheese;
  int p1, p2, mud1, mud2;
  float s1, s2, progress;
  const uint8_t* maze;
  int spatial;
};
__device__ __forceinline__ RowView row_view(const EvalRow& r, const ar_game_pod* games) {
  RowView v;
  const ar_game_pod& g = games[r.game_idx];
  v.cheese = r.cheese;
  v.p1 = r.pos & 0xff; v.p2 = (r.pos >> 8) & 0xff; v.mud1 = (r.pos >> 16) & 0xff; v.mud2 = r.pos >> 24;
  v.s1 = 0.5f * (float)(r.score & 0xffff);
  v.s2 = 0.5f * (float)(r.score >> 16);
  v.progress = r.max_turns > 0 ? (float)r.turn / (float)r.max_turns : 0.0f;
  v.maze = g.move_cost;
  v.spatial = (int)g.width * g.height;
  return v;
}
// exact f32 observation element k
__device__ __forceinline__ float obs_elem(const RowView& v, int k) {
  const int S = v.spatial;
  if (k < 4 * S) {
    int c = v.maze[k];
    return c == 0 ? -1.0f : (c >= 2 ? (float)c / 10.0f : 1.0f / 10.0f);
  }
  k -= 4 * S;
  if (k < S) return k == v.p1 ? 1.0f : 0.0f;
  k -= S;
  if (k < S) return k == v.p2 ? 1.0f : 0.0f;
  k -= S;
  if (k < S) return ((v.cheese >> k) & 1ull) ? 1.0f : 0.0f;
  k -= S;
  switch (k) {
    case 0: return v.s1 - v.s2;
    case 1: return v.progress;
    case 2: return (float)v.mud1 / 10.0f;
    case 3: return (float)v.mud2 / 10.0f;
    case 4: return v.s1 / 10.0f;
    case 5: return v.s2 / 10.0f;
    default: return 0.0f;
  }
}

__global__ void encode_f32_kernel(const EvalRow* rows, int n, const ar_game_pod* games, int obs_dim,
                                  float* out) {
  int i = blockIdx.x;
  if (i >= n) return;
  RowView v = row_view(rows[i], games);
  for (int k = threadIdx.x; k < obs_dim; k += blockDim.x) out[(size_t)i * obs_dim + k] = obs_elem(v, k);
}

// ---------------------------------------------------------------------------------------
// Fused encode + MLP forward
// ---------------------------------------------------------------------------------------
constexpr int TILE_M = 128;
constexpr int KB = 64;                // K elements per operand block (128 B of bf16)
constexpr int A_BLOCK_BYTES = TILE_M * KB * 2;  // 16 KB
constexpr int W_STAGE_BYTES = 256 * KB * 2;     // 32 KB
constexpr int N_STAGES = 3;
constexpr int MLP_THREADS = 192;

struct MlpSmem {
  uint64_t w_full[N_STAGES];
  uint64_t w_empty[N_STAGES];
  uint64_t a_ready;
  uint64_t mma_done;
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// One CTA = 128 rows per tile.  warps 0-3: encode + epilogue (thread t owns row t and TMEM lane
// t), warp 4: TMA weight producer, warp 5: MMA issuer (and TMEM allocation).
__global__ void __launch_bounds__(MLP_THREADS, 1)
mlp_forward_kernel(const EvalRow* __restrict__ rows, const uint32_t* __restrict__ n_rows_ptr, int n_rows_arg,
                   const ar_game_pod* __restrict__ games, MlpWeights w, float* __restrict__ out,
                   int* __restrict__ error_flag) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* xa = smem;                                  // A operand / hidden activations
  uint8_t* ws = smem + w.k1_blocks * A_BLOCK_BYTES;    // weight stages
  MlpSmem* sh = reinterpret_cast<MlpSmem*>(ws + N_STAGES * W_STAGE_BYTES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_rows = n_rows_ptr ? (int)*n_rows_ptr : n_rows_arg;
  const int n_tiles = (n_rows + TILE_M - 1) / TILE_M;
  if ((int)blockIdx.x >= n_tiles) return;

  if (tid == 0) {
    for (int s = 0; s < N_STAGES; ++s) {
      mbar_init(&sh->w_full[s], 1);
      mbar_init(&sh->w_empty[s], 1);
    }
    mbar_init(&sh->a_ready, 128);
    mbar_init(&sh->mma_done, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(&sh->tmem_base, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh->tmem_base;

  const int k1b = w.k1_blocks;  // layer-1 K blocks (obs_dim padded to 64)
  // chunk schedule per tile: k1b blocks of W1, 4 of W2, 4 of W3
  const int chunks_per_tile = k1b + 8;

  if (warp == 4) {
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        for (int c = 0; c < chunks_per_tile; ++c, ++it) {
          int s = it % N_STAGES;
          uint32_t ph = (it / N_STAGES) & 1;
          mbar_wait(&sh->w_empty[s], ph ^ 1);
          const uint8_t* src;
          uint32_t bytes;
          if (c < k1b) { src = w.w1 + (size_t)c * W_STAGE_BYTES; bytes = W_STAGE_BYTES; }
          else if (c < k1b + 4) { src = w.w2 + (size_t)(c - k1b) * W_STAGE_BYTES; bytes = W_STAGE_BYTES; }
          else { src = w.w3 + (size_t)(c - k1b - 4) * (16 * KB * 2); bytes = 16 * KB * 2; }
          mbar_expect_tx(&sh->w_full[s], bytes);
          tma_bulk_g2s(ws + s * W_STAGE_BYTES, src, bytes, &sh->w_full[s]);
        }
      }
    }
  } else if (warp == 5) {
    // ================= MMA issuer =================
    if (lane == 0) {
      uint32_t it = 0, a_phase = 0;
      const uint32_t idesc256 = umma_idesc(TILE_M, 256), idesc16 = umma_idesc(TILE_M, 16);
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        for (int layer = 0; layer < 3; ++layer) {
          int nkb = layer == 0 ? k1b : 4;
          uint32_t idesc = layer == 2 ? idesc16 : idesc256;
          mbar_wait(&sh->a_ready, a_phase);
          a_phase ^= 1;
          tc_fence_after();
          for (int kb = 0; kb < nkb; ++kb, ++it) {
            int s = it % N_STAGES;
            uint32_t ph = (it / N_STAGES) & 1;
            mbar_wait(&sh->w_full[s], ph);
            tc_fence_after();
            uint64_t da = umma_desc_sw128(smem_u32(xa + kb * A_BLOCK_BYTES));
            uint64_t db = umma_desc_sw128(smem_u32(ws + s * W_STAGE_BYTES));
#pragma unroll
            for (int k = 0; k < KB / 16; ++k)  // UMMA_K = 16 bf16 = 32 B = 2 descriptor units
              umma_bf16(tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
            umma_commit(&sh->w_empty[s]);  // frees the weight stage when these MMAs retire
          }
          umma_commit(&sh->mma_done);
        }
      }
    }
  } else {
    // ================= encode + epilogue (128 threads, thread = row) =================
    const int r = tid;  // row in tile, TMEM lane
    const uint32_t t_lane = tmem + ((uint32_t)(warp * 32) << 16);
    uint32_t done_phase = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int row = t * TILE_M + r;
      const bool live = row < n_rows;
      // ---- encode observation -> bf16 A operand (k1b blocks of [128 x 64], SW128)
      {
        RowView v;
        if (live) v = row_view(rows[row], games);
        for (int kb = 0; kb < k1b; ++kb) {
#pragma unroll 2
          for (int cc = 0; cc < 8; ++cc) {
            float e[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) e[j] = live ? obs_elem(v, kb * KB + cc * 8 + j) : 0.0f;
            uint4 pk = make_uint4(pack_bf16(e[0], e[1]), pack_bf16(e[2], e[3]), pack_bf16(e[4], e[5]),
                                  pack_bf16(e[6], e[7]));
            *reinterpret_cast<uint4*>(xa + kb * A_BLOCK_BYTES + sw128_offset(r, cc * 8)) = pk;
          }
        }
      }
      fence_proxy_async();
      mbar_arrive(&sh->a_ready);
      // ---- hidden layers: D -> +bias, ReLU -> bf16 A operand of the next layer
      for (int layer = 0; layer < 2; ++layer) {
        const float* bias = layer == 0 ? w.b1 : w.b2;
        mbar_wait(&sh->mma_done, done_phase);
        done_phase ^= 1;
        tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < 256; c0 += 16) {
          float v[16];
          tmem_ld16(t_lane + c0, v);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j] + __ldg(bias + c0 + j), 0.0f);
          int kb = c0 >> 6, col = c0 & 63;
          uint4 p0 = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]),
                                pack_bf16(v[6], v[7]));
          uint4 p1 = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]),
                                pack_bf16(v[14], v[15]));
          *reinterpret_cast<uint4*>(xa + kb * A_BLOCK_BYTES + sw128_offset(r, col)) = p0;
          *reinterpret_cast<uint4*>(xa + kb * A_BLOCK_BYTES + sw128_offset(r, col + 8)) = p1;
        }
        tc_fence_before();
        fence_proxy_async();
        mbar_arrive(&sh->a_ready);
      }
      // ---- heads: logits p1[0:5], p2[5:10], value[10:12]
      mbar_wait(&sh->mma_done, done_phase);
      done_phase ^= 1;
      tc_fence_after();
      float z[16];
      tmem_ld16(t_lane, z);
      tc_fence_before();
      if (live) {
#pragma unroll
        for (int j = 0; j < 12; ++j) z[j] += __ldg(w.b3 + j);
        float o[12];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float m = z[5 * h];
#pragma unroll
          for (int j = 1; j < 5; ++j) m = fmaxf(m, z[5 * h + j]);
          float e[5], s = 0.0f;
#pragma unroll
          for (int j = 0; j < 5; ++j) { e[j] = expf(z[5 * h + j] - m); s += e[j]; }
#pragma unroll
          for (int j = 0; j < 5; ++j) o[5 * h + j] = e[j] / s;
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          float x = z[10 + j];
          o[10 + j] = x > 20.0f ? x : log1pf(expf(x));  // F.softplus (beta 1, threshold 20)
        }
        bool ok = true;
#pragma unroll
        for (int j = 0; j < 12; ++j) ok = ok && isfinite(o[j]);
        if (!ok) atomicCAS(error_flag, 0, (int)AR_ERR_NONFINITE);  // onnx.rs:233-241
        float4* dst = reinterpret_cast<float4*>(out + (size_t)row * 12);
        dst[0] = make_float4(o[0], o[1], o[2], o[3]);
        dst[1] = make_float4(o[4], o[5], o[6], o[7]);
        dst[2] = make_float4(o[8], o[9], o[10], o[11]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem, 256);
}

// ---------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------
static uint16_t f32_to_bf16(float f) {  // round to nearest even
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  uint32_t lsb = (u >> 16) & 1u;
  u += 0x7fffu + lsb;
  return (uint16_t)(u >> 16);
}

static const ar_tensor_desc* find_tensor(const ar_tensor_desc* t, int n, const char* name) {
  for (int i = 0; i < n; ++i)
    if (t[i].name && strcmp(t[i].name, name) == 0) return &t[i];
  return nullptr;
}

// [N x K] row-major f32 (already BN-folded) -> K/64 blocks of [n_pad x 64] bf16, SW128 image
static std::vector<uint8_t> swizzled_image(const std::vector<float>& W, int N, int K, int n_pad, int k_blocks) {
  std::vector<uint8_t> img((size_t)k_blocks * n_pad * KB * 2, 0);
  for (int kb = 0; kb < k_blocks; ++kb)
    for (int n = 0; n < N; ++n)
      for (int c = 0; c < KB; ++c) {
        int k = kb * KB + c;
        float v = k < K ? W[(size_t)n * K + k] : 0.0f;
        uint16_t h = f32_to_bf16(v);
        size_t off = (size_t)kb * n_pad * KB * 2 + sw128_offset(n, c);
        memcpy(&img[off], &h, 2);
      }
  return img;
}

void MlpModel::release() {
  cudaFree(d_w1); cudaFree(d_w2); cudaFree(d_w3); cudaFree(d_b);
  d_w1 = d_w2 = d_w3 = nullptr;
  d_b = nullptr;
  loaded = false;
}

// Fold eval-mode BatchNorm1d (running stats, eps 1e-5) into the preceding Linear:
//   y = gamma * (Wx + b - mean) / sqrt(var + eps) + beta
static bool fold_linear_bn(const ar_tensor_desc* t, int n, const std::string& lin, const std::string& bn,
                           int out_f, int in_f, std::vector<float>& W, std::vector<float>& b, std::string& err) {
  auto get = [&](const std::string& name, int64_t numel) -> const float* {
    const ar_tensor_desc* d = find_tensor(t, n, name.c_str());
    if (!d) { err = "missing tensor " + name; return nullptr; }
    int64_t ne = 1;
    for (int i = 0; i < d->ndim; ++i) ne *= d->shape[i];
    if (ne != numel) { err = "tensor " + name + " has " + std::to_string(ne) + " elements, expected " + std::to_string(numel); return nullptr; }
    return d->data;
  };
  const float* w = get(lin + ".weight", (int64_t)out_f * in_f);
  const float* bi = get(lin + ".bias", out_f);
  if (!w || !bi) return false;
  W.assign(w, w + (size_t)out_f * in_f);
  b.assign(bi, bi + out_f);
  if (!bn.empty()) {
    const float* g = get(bn + ".weight", out_f);
    const float* be = get(bn + ".bias", out_f);
    const float* mu = get(bn + ".running_mean", out_f);
    const float* var = get(bn + ".running_var", out_f);
    if (!g || !be || !mu || !var) return false;
    for (int o = 0; o < out_f; ++o) {
      float s = g[o] / sqrtf(var[o] + 1e-5f);
      for (int i = 0; i < in_f; ++i) W[(size_t)o * in_f + i] *= s;
      b[o] = (b[o] - mu[o]) * s + be[o];
    }
  }
  return true;
}

int MlpModel::load(const ar_tensor_desc* t, int n, int width, int height, std::string& err) {
  release();
  obs_dim = 7 * width * height + 6;
  const ar_tensor_desc* w1d = find_tensor(t, n, "trunk.0.weight");
  if (!w1d || w1d->ndim != 2) { err = "trunk.0.weight missing"; return AR_ERR_INVALID_ARG; }
  if (w1d->shape[0] != 256) { err = "only hidden_dim = 256 is supported by the fused MLP kernel"; return AR_ERR_UNSUPPORTED; }
  if (w1d->shape[1] != obs_dim) { err = "trunk.0.weight does not match obs_dim " + std::to_string(obs_dim); return AR_ERR_INVALID_ARG; }
  k1_blocks = (obs_dim + KB - 1) / KB;
  if (k1_blocks > 6) { err = "obs_dim > 384 does not fit the shared-memory budget of this kernel"; return AR_ERR_UNSUPPORTED; }
  std::vector<float> W1, b1, W2, b2, Wp1, bp1, Wp2, bp2, Wv, bv;
  if (!fold_linear_bn(t, n, "trunk.0", "trunk.1", 256, obs_dim, W1, b1, err)) return AR_ERR_INVALID_ARG;
  if (!fold_linear_bn(t, n, "trunk.4", "trunk.5", 256, 256, W2, b2, err)) return AR_ERR_INVALID_ARG;
  if (!fold_linear_bn(t, n, "policy_p1_head", "", 5, 256, Wp1, bp1, err)) return AR_ERR_INVALID_ARG;
  if (!fold_linear_bn(t, n, "policy_p2_head", "", 5, 256, Wp2, bp2, err)) return AR_ERR_INVALID_ARG;
  if (!fold_linear_bn(t, n, "value_head", "", 2, 256, Wv, bv, err)) return AR_ERR_INVALID_ARG;
  std::vector<float> W3((size_t)16 * 256, 0.0f), b3(16, 0.0f);
  memcpy(&W3[0], Wp1.data(), 5 * 256 * 4);
  memcpy(&W3[5 * 256], Wp2.data(), 5 * 256 * 4);
  memcpy(&W3[10 * 256], Wv.data(), 2 * 256 * 4);
  for (int i = 0; i < 5; ++i) { b3[i] = bp1[i]; b3[5 + i] = bp2[i]; }
  b3[10] = bv[0]; b3[11] = bv[1];
  std::vector<uint8_t> i1 = swizzled_image(W1, 256, obs_dim, 256, k1_blocks);
  std::vector<uint8_t> i2 = swizzled_image(W2, 256, 256, 256, 4);
  std::vector<uint8_t> i3 = swizzled_image(W3, 16, 256, 16, 4);
  std::vector<float> ball(256 + 256 + 16);
  memcpy(&ball[0], b1.data(), 1024); memcpy(&ball[256], b2.data(), 1024); memcpy(&ball[512], b3.data(), 64);
#define CKN(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(_e); return AR_ERR_CUDA; } } while (0)
  CKN(cudaMalloc(&d_w1, i1.size())); CKN(cudaMalloc(&d_w2, i2.size())); CKN(cudaMalloc(&d_w3, i3.size()));
  CKN(cudaMalloc(&d_b, ball.size() * 4));
  CKN(cudaMemcpy(d_w1, i1.data(), i1.size(), cudaMemcpyHostToDevice));
  CKN(cudaMemcpy(d_w2, i2.data(), i2.size(), cudaMemcpyHostToDevice));
  CKN(cudaMemcpy(d_w3, i3.data(), i3.size(), cudaMemcpyHostToDevice));
  CKN(cudaMemcpy(d_b, ball.data(), ball.size() * 4, cudaMemcpyHostToDevice));
  smem_bytes = (size_t)k1_blocks * A_BLOCK_BYTES + N_STAGES * W_STAGE_BYTES + sizeof(MlpSmem) + 1024;
  CKN(cudaFuncSetAttribute(mlp_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  int dev = 0;
  CKN(cudaGetDevice(&dev));
  CKN(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
#undef CKN
  loaded = true;
  return AR_OK;
}

// n_rows_dev != nullptr: the row count is read on device (the self-play loop never syncs).
cudaError_t MlpModel::forward(const EvalRow* rows, const uint32_t* n_rows_dev, int n_rows_max,
                              const ar_game_pod* games, float* out, int* error_flag, cudaStream_t stream) const {
  if (n_rows_max <= 0) return cudaSuccess;
  MlpWeights w;
  w.w1 = d_w1; w.w2 = d_w2; w.w3 = d_w3;
  w.b1 = d_b; w.b2 = d_b + 256; w.b3 = d_b + 512;
  w.k1_blocks = k1_blocks;
  int tiles = (n_rows_max + TILE_M - 1) / TILE_M;
  int grid = tiles < n_sms ? tiles : n_sms;
  mlp_forward_kernel<<<grid, MLP_THREADS, smem_bytes, stream>>>(rows, n_rows_dev, n_rows_max, games, w, out,
                                                                 error_flag);
  return cudaGetLastError();
}

cudaError_t encode_f32(const EvalRow* rows, int n, const ar_game_pod* games, int obs_dim, float* out,
                       cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  encode_f32_kernel<<<n, 128, 0, stream>>>(rows, n, games, obs_dim, out);
  return cudaGetLastError();
}

}  // namespace ar
