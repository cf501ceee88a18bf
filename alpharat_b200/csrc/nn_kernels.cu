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

#include <algorithm>

#include "nn_common.cuh"

namespace ar {

__global__ void maze_table_kernel(const ar_game_pod* __restrict__ games, int n, uint16_t* __restrict__ out) {
  const int g = blockIdx.x, k = threadIdx.x;
  if (g >= n) return;
  const ar_game_pod& pod = games[g];
  const int cells = (int)pod.width * pod.height;
  float v = 0.0f;
  if (k < cells * 4) {
    const int c = pod.move_cost[k];
    v = c == 0 ? -1.0f : (c >= 2 ? (float)c / 10.0f : 1.0f / 10.0f);
  }
  out[(size_t)g * MAZE_TAB_STRIDE + k] = k < cells * 4 ? bf16_bits(v) : (uint16_t)0;
}

cudaError_t build_maze_table(const ar_game_pod* games, int n, uint16_t* out, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  maze_table_kernel<<<n, MAZE_TAB_STRIDE, 0, stream>>>(games, n, out);
  return cudaGetLastError();
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
constexpr int A_BLOCK_BYTES = TILE_M * KB * 2;  // 16 KB
constexpr int W_STAGE_BYTES = 256 * KB * 2;     // 32 KB
constexpr int N_STAGES = 3;
constexpr int EPI_WARPS = 8;                    // two threads per tile row, 128 output columns each
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int MLP_THREADS = EPI_THREADS + 64;   // + TMA warp + MMA warp

struct MlpSmem {
  uint64_t w_full[N_STAGES];
  uint64_t w_empty[N_STAGES];
  uint64_t a_ready;
  uint64_t mma_done;
  uint32_t tmem_base;
};

// 16 TMEM columns without waiting: several loads are in flight before one wait
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t r[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory"); }

// One CTA = 128 rows per tile.  warps 0-7: encode + epilogue (threads r and r + 128 own row r = TMEM lane r,
// 128 of the 256 hidden columns each), warp 8: TMA weight producer, warp 9: MMA issuer (stays converged so that
// the descriptors live in uniform registers; tcgen05 instructions predicated on lane 0) and TMEM allocation.
__global__ void __launch_bounds__(MLP_THREADS, 1)
mlp_forward_kernel(const EvalRow* __restrict__ rows, const uint32_t* __restrict__ n_rows_ptr, int n_rows_arg,
                   const ar_game_pod* __restrict__ games, const uint16_t* __restrict__ maze_tab, MlpWeights w,
                   float* __restrict__ out, int* __restrict__ error_flag) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* xa = smem;                                  // A operand / hidden activations
  // the buffer holds the encoded observation (k1_blocks K-blocks) and later the 256 hidden activations (4)
  uint8_t* ws = smem + (w.k1_blocks > 4 ? w.k1_blocks : 4) * A_BLOCK_BYTES;    // weight stages
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
    mbar_init(&sh->a_ready, EPI_WARPS);
    mbar_init(&sh->mma_done, 1);
    fence_barrier_init();
  }
  if (warp == EPI_WARPS + 1) tmem_alloc(&sh->tmem_base, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh->tmem_base;

  const int k1b = w.k1_blocks;  // layer-1 K blocks (obs_dim padded to 64)
  // chunk schedule per tile: k1b blocks of W1, 4 of W2, 4 of W3
  const int chunks_per_tile = k1b + 8;

  if (warp == EPI_WARPS) {
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
  } else if (warp == EPI_WARPS + 1) {
    // ================= MMA issuer (whole warp converged) =================
    const uint32_t issue = lane == 0 ? 1u : 0u;
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    uint32_t it = 0, a_phase = 0;
    const uint32_t idesc256 = umma_idesc(TILE_M, 256), idesc16 = umma_idesc(TILE_M, 16);
    const uint64_t a_desc0 = umma_desc_sw128(smem_u32(xa)), w_desc0 = umma_desc_sw128(smem_u32(ws));
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      for (int layer = 0; layer < 3; ++layer) {
        const int nkb = layer == 0 ? k1b : 4;
        const uint32_t idesc = layer == 2 ? idesc16 : idesc256;
        mbar_wait_warp(&sh->a_ready, a_phase);
        a_phase ^= 1;
        tc_fence_after();
#pragma unroll 1
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = __shfl_sync(0xffffffffu, it % N_STAGES, 0);
          const uint32_t ph = (it / N_STAGES) & 1;
          mbar_wait_warp(&sh->w_full[s], ph);
          const uint64_t da = a_desc0 + (uint64_t)(kb * (A_BLOCK_BYTES >> 4));
          const uint64_t db = w_desc0 + (uint64_t)(s * (W_STAGE_BYTES >> 4));
          umma_bf16_pred(tmem_u, da, db, idesc, kb ? 1u : 0u, issue);
#pragma unroll
          for (int k = 1; k < KB / 16; ++k)  // UMMA_K = 16 bf16 = 32 B = 2 descriptor units
            umma_bf16_pred(tmem_u, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, 1u, issue);
          umma_commit_pred(&sh->w_empty[s], issue);  // frees the weight stage when these MMAs retire
        }
        umma_commit_pred(&sh->mma_done, issue);
      }
    }
    __syncwarp();
  } else {
    // ================= encode + epilogue (256 threads, two per row) =================
    const int r = tid & (TILE_M - 1), half = tid >> 7;  // row in tile = TMEM lane; column half
    const uint32_t t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t done_phase = 0;
    // every thread orders its generic-proxy stores before the async proxy, one lane per warp arrives
    auto signal_a = [&]() {
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->a_ready);
    };
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int row = t * TILE_M + r;
      const bool live = row < n_rows;
      // ---- encode observation -> bf16 A operand (k1b blocks of [128 x 64], SW128): the maze
      //      channels are copied from the per-game table, the rest of the row is zeroed (each thread of the
      //      pair fills half of the row) and the few non-zero elements (two positions, the cheese, six
      //      scalars) are stored one by one after a barrier
      {
        RowView v;
        const uint4* mt = nullptr;
        int n_maze_pieces = 0;
        if (live) {
          const EvalRow er = rows[row];
          v = row_view(er, games);
          mt = reinterpret_cast<const uint4*>(maze_tab + (size_t)er.game_idx * MAZE_TAB_STRIDE);
          n_maze_pieces = (4 * v.spatial + 7) >> 3;
        }
        const int p_half = k1b * 4;
        for (int p = half * p_half; p < (half + 1) * p_half; ++p) {
          uint4 pk = make_uint4(0, 0, 0, 0);
          if (p < n_maze_pieces) pk = __ldg(mt + p);
          *reinterpret_cast<uint4*>(xa + (p >> 3) * A_BLOCK_BYTES + sw128_offset(r, (p & 7) * 8)) = pk;
        }
        epi_barrier();  // order the 16-byte fills (both threads of a row) before the element stores below
        if (live && half == 0) {
          const int S = v.spatial;
          put_elem(xa, A_BLOCK_BYTES, r, 4 * S + v.p1, BF16_ONE);
          put_elem(xa, A_BLOCK_BYTES, r, 5 * S + v.p2, BF16_ONE);
          for (uint64_t c = v.cheese; c; c &= c - 1) put_elem(xa, A_BLOCK_BYTES, r, 6 * S + (__ffsll((long long)c) - 1), BF16_ONE);
#pragma unroll
          for (int j = 0; j < 6; ++j) put_elem(xa, A_BLOCK_BYTES, r, 7 * S + j, bf16_bits(obs_elem(v, 7 * S + j)));
        }
      }
      signal_a();
      // ---- hidden layers: D -> +bias, ReLU -> bf16 A operand of the next layer
      for (int layer = 0; layer < 2; ++layer) {
        const float* bias = layer == 0 ? w.b1 : w.b2;
        mbar_wait(&sh->mma_done, done_phase);
        done_phase ^= 1;
        tc_fence_after();
#pragma unroll 1
        for (int c64 = half * 128; c64 < half * 128 + 128; c64 += 64) {
          uint32_t acc[64];  // four loads in flight, one wait
#pragma unroll
          for (int cc = 0; cc < 64; cc += 16) tmem_ld16_nowait(t_lane + c64 + cc, acc + cc);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int cc = 0; cc < 64; cc += 16) {
            const int c0 = c64 + cc;
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + c0 + j));
              v[j] = fmaxf(__uint_as_float(acc[cc + j]) + b4.x, 0.0f);
              v[j + 1] = fmaxf(__uint_as_float(acc[cc + j + 1]) + b4.y, 0.0f);
              v[j + 2] = fmaxf(__uint_as_float(acc[cc + j + 2]) + b4.z, 0.0f);
              v[j + 3] = fmaxf(__uint_as_float(acc[cc + j + 3]) + b4.w, 0.0f);
            }
            const int kb = c0 >> 6, col = c0 & 63;
            uint4 p0 = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]),
                                  pack_bf16(v[6], v[7]));
            uint4 p1 = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]),
                                  pack_bf16(v[14], v[15]));
            *reinterpret_cast<uint4*>(xa + kb * A_BLOCK_BYTES + sw128_offset(r, col)) = p0;
            *reinterpret_cast<uint4*>(xa + kb * A_BLOCK_BYTES + sw128_offset(r, col + 8)) = p1;
          }
        }
        signal_a();
      }
      // ---- heads: logits p1[0:5], p2[5:10], value[10:12]
      mbar_wait(&sh->mma_done, done_phase);
      done_phase ^= 1;
      tc_fence_after();
      float z[16];
      tmem_ld16(t_lane, z);
      tc_fence_before();
      if (live && half == 0) {
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
  if (warp == EPI_WARPS + 1) tmem_dealloc(tmem, 256);
}

// ---------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------
void MlpModel::release() {
  cudaFree(d_w1); cudaFree(d_w2); cudaFree(d_w3); cudaFree(d_b);
  d_w1 = d_w2 = d_w3 = nullptr;
  d_b = nullptr;
  loaded = false;
}

int MlpModel::load(const ar_tensor_desc* t, int n, int width, int height, std::string& err) {
  release();
  obs_dim = 7 * width * height + 6;
  const ar_tensor_desc* w1d = find_tensor(t, n, "trunk.0.weight");
  if (!w1d || w1d->ndim != 2) { err = "trunk.0.weight missing"; return AR_ERR_INVALID_ARG; }
  // The kernel computes 256 hidden columns.  A narrower trunk (hidden_dim < 256) runs zero-padded: the extra units
  // have zero weights and zero bias, so they are exactly 0 after the ReLU and add exactly 0 downstream.
  const int hid = (int)w1d->shape[0];
  if (hid < 1 || hid > 256) { err = "hidden_dim must be in [1, 256] for the fused MLP kernel (narrower trunks are zero-padded to its 256 columns), got " + std::to_string(hid); return AR_ERR_UNSUPPORTED; }
  if (w1d->shape[1] != obs_dim) { err = "trunk.0.weight does not match obs_dim " + std::to_string(obs_dim); return AR_ERR_INVALID_ARG; }
  k1_blocks = (obs_dim + KB - 1) / KB;
  if (k1_blocks > 6) { err = "obs_dim > 384 does not fit the shared-memory budget of this kernel"; return AR_ERR_UNSUPPORTED; }
  std::vector<float> W1h, b1h, W2h, b2h, Wp1, bp1, Wp2, bp2, Wv, bv;
  if (!fold_linear_bn(t, n, "trunk.0", "trunk.1", hid, obs_dim, W1h, b1h, err)) return AR_ERR_INVALID_ARG;
  if (!fold_linear_bn(t, n, "trunk.4", "trunk.5", hid, hid, W2h, b2h, err)) return AR_ERR_INVALID_ARG;
  if (!fold_linear_bn(t, n, "policy_p1_head", "", 5, hid, Wp1, bp1, err)) return AR_ERR_INVALID_ARG;
  if (!fold_linear_bn(t, n, "policy_p2_head", "", 5, hid, Wp2, bp2, err)) return AR_ERR_INVALID_ARG;
  if (!fold_linear_bn(t, n, "value_head", "", 2, hid, Wv, bv, err)) return AR_ERR_INVALID_ARG;
  std::vector<float> W1((size_t)256 * obs_dim, 0.0f), b1(256, 0.0f), W2((size_t)256 * 256, 0.0f), b2(256, 0.0f);
  std::vector<float> W3((size_t)16 * 256, 0.0f), b3(16, 0.0f);
  memcpy(W1.data(), W1h.data(), (size_t)hid * obs_dim * 4);
  memcpy(b1.data(), b1h.data(), (size_t)hid * 4);
  memcpy(b2.data(), b2h.data(), (size_t)hid * 4);
  for (int r = 0; r < hid; ++r) memcpy(&W2[(size_t)r * 256], &W2h[(size_t)r * hid], (size_t)hid * 4);
  for (int j = 0; j < 5; ++j) {
    memcpy(&W3[(size_t)j * 256], &Wp1[(size_t)j * hid], (size_t)hid * 4);
    memcpy(&W3[(size_t)(5 + j) * 256], &Wp2[(size_t)j * hid], (size_t)hid * 4);
  }
  for (int j = 0; j < 2; ++j) memcpy(&W3[(size_t)(10 + j) * 256], &Wv[(size_t)j * hid], (size_t)hid * 4);
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
  smem_bytes = (size_t)std::max(k1_blocks, 4) * A_BLOCK_BYTES + N_STAGES * W_STAGE_BYTES + sizeof(MlpSmem) + 1024;
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
                              const ar_game_pod* games, const uint16_t* maze_tab, float* out, int* error_flag,
                              cudaStream_t stream) const {
  if (n_rows_max <= 0) return cudaSuccess;
  MlpWeights w;
  w.w1 = d_w1; w.w2 = d_w2; w.w3 = d_w3;
  w.b1 = d_b; w.b2 = d_b + 256; w.b3 = d_b + 512;
  w.k1_blocks = k1_blocks;
  int tiles = (n_rows_max + TILE_M - 1) / TILE_M;
  int grid = tiles < n_sms ? tiles : n_sms;
  mlp_forward_kernel<<<grid, MLP_THREADS, smem_bytes, stream>>>(rows, n_rows_dev, n_rows_max, games, maze_tab, w,
                                                                 out, error_flag);
  return cudaGetLastError();
}

cudaError_t encode_f32(const EvalRow* rows, int n, const ar_game_pod* games, int obs_dim, float* out,
                       cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  encode_f32_kernel<<<n, 128, 0, stream>>>(rows, n, games, obs_dim, out);
  return cudaGetLastError();
}

}  // namespace ar
