// nn_common.cuh — shared by the leaf-evaluator kernels (nn_kernels.cu, nn_symmetric.cu, nn_cnn.cu):
// tcgen05 / TMA / mbarrier PTX wrappers, the 128B-swizzled K-major operand layout, the on-device
// observation encoder (flat_encoder.rs:52-124) and the host-side weight packing helpers.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "nn_api.cuh"

namespace ar {

constexpr int KB = 64;  // K elements per operand block (128 B of bf16)

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
// mbar_wait for a warp that must stay converged: the loop exit is a warp vote, so the compiler keeps
// treating the code after it as warp-uniform.
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity) {
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
  } while (!__all_sync(0xffffffffu, done));
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
// Predicated forms for an issuer warp that stays converged (operands warp-uniform): only lanes with
// `issue` != 0 execute the tcgen05 instruction.
__device__ __forceinline__ void umma_bf16_pred(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate, uint32_t issue) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(issue)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pred(uint64_t* bar, uint32_t issue) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "setp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(issue)
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


constexpr uint16_t BF16_ONE = 0x3f80;
__device__ __forceinline__ uint16_t bf16_bits(float x) {
  __nv_bfloat16 h = __float2bfloat16_rn(x);
  return *reinterpret_cast<uint16_t*>(&h);
}
// store one bf16 element (row r, column k) of a K-major SW128 operand made of [128 x 64] blocks
__device__ __forceinline__ void put_elem(uint8_t* a_blocks, int block_bytes, int r, int k, uint16_t bits) {
  *reinterpret_cast<uint16_t*>(a_blocks + (k >> 6) * block_bytes + sw128_offset(r, k & 63)) = bits;
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}


// ---------------------------------------------------------------------------------------
// Host-side packing helpers
// ---------------------------------------------------------------------------------------
static inline uint16_t f32_to_bf16(float f) {  // round to nearest even
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  uint32_t lsb = (u >> 16) & 1u;
  u += 0x7fffu + lsb;
  return (uint16_t)(u >> 16);
}

static inline const ar_tensor_desc* find_tensor(const ar_tensor_desc* t, int n, const char* name) {
  for (int i = 0; i < n; ++i)
    if (t[i].name && strcmp(t[i].name, name) == 0) return &t[i];
  return nullptr;
}

// [N x K] row-major f32 (already BN-folded) -> K/64 blocks of [n_pad x 64] bf16, SW128 image
static inline std::vector<uint8_t> swizzled_image(const std::vector<float>& W, int N, int K, int n_pad, int k_blocks) {
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

// Fold eval-mode BatchNorm1d (running stats, eps 1e-5) into the preceding Linear:
//   y = gamma * (Wx + b - mean) / sqrt(var + eps) + beta
static inline bool fold_linear_bn(const ar_tensor_desc* t, int n, const std::string& lin, const std::string& bn,
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

}  // namespace ar
