// nn_cnn.cu — PyRatCNN (residual trunk with optional global-pooling blocks + DeepSet heads) leaf
// evaluator on tcgen05 tensor cores.
//
// Graph: PyRatCNN.predict, alpharat/nn/models/cnn/model.py:117-230 (eval mode), with
//   stem     Conv3x3(5 -> C, no bias) -> BN -> ReLU                      model.py:181
//   ResBlock       x + conv2(ReLU(BN2(conv1(ReLU(BN1(x))))))             cnn/blocks.py:10-30
//   GPoolResBlock  ... + pool_linear([mean, max](pool_conv(ReLU(pool_bn(x)))))   cnn/blocks.py:33-79
//   f_i = trunk features at player i's cell; e_i = ReLU(Linear(3, P)([score, mud, progress]))
//   h_i = ReLU(Linear(C+P, H)(cat(f_i, e_i))); agg = h_1 + h_2
//   MLPPolicyHead / PointValueHead on cat(h_i, agg)                      cnn/heads.py:10-40
// Supported: C = 64 trunk channels, any sequence of res / gpool blocks (gpool_channels 16 or 32),
// P <= 32, H <= 64, `mlp` policy head and `point` value head (configs/model/cnn*.yaml).
//
// B200 mapping: a 3x3 convolution is an implicit GEMM with M = 128 spatial rows (one row per TMEM
// lane), N = 64, K = 9 taps x 64 channels, and the im2col image is never materialised.  A position's
// board is laid out with one zero column and one zero row of padding ((w+1) x (h+1) rows), so the
// input of tap (dy, dx) for tile row r is simply row r + dy * (w+1) + dx of the same activation buffer:
// every tap is the same 128B-swizzled K-major operand read through a shared-memory descriptor whose
// start address is shifted by whole rows (the swizzle is a function of the shared-memory address, and
// the buffer is written with the same address-based swizzle).  An epilogue thread therefore stores
// its 128-byte activation vector once instead of nine times, and out-of-board taps read the padding.
// A pipeline (one 128-row tile at a time: encode -> MMA -> epilogue -> MMA ...) is a serial chain, so two
// CTAs share an SM (NG = 1: 113 KB of shared memory and 256 TMEM columns each) and one CTA's epilogue
// overlaps the other's MMAs; a pipeline has 8 epilogue warps (two threads per tile row, 32 channels each),
// a TMA warp streaming 8 KB weight taps through an 8-stage bulk-copy ring and an MMA warp that stays
// converged (descriptors in uniform registers, tcgen05 instructions predicated on lane 0).
// The residual stream never leaves TMEM: conv2 accumulates straight onto it (accumulate = 1 from
// the first MMA), conv1 goes to a second accumulator, the gpool 1x1 convolution to a third.
// BatchNorms that follow a convolution are folded into its weights; pre-activation BatchNorms are
// applied in the epilogue.  The tiny per-player head (3 -> P -> H -> 6) runs on CUDA cores.
#include "nn_common.cuh"

namespace ar {
namespace cnn {

constexpr int TILE_M = 128;
constexpr int C = 64;
#ifndef AR_CNN_NG
#define AR_CNN_NG 1
#endif
constexpr int NG = AR_CNN_NG;                    // independent 128-row pipelines per CTA (1: two CTAs share an SM)
constexpr int CTAS_PER_SM = NG == 1 ? 2 : 1;
constexpr int GT = 2 * TILE_M;                   // epilogue threads of one pipeline: two per tile row (32 channels each)
constexpr int EPI_THREADS = GT * NG;
constexpr int THREADS = EPI_THREADS + 64;        // + TMA warp + MMA warp
constexpr int GUARD = 16;                        // zero rows before / after a group's activation rows
constexpr int A_BYTES = (TILE_M + 2 * GUARD) * 128;  // 20 KB activation buffer of one group
constexpr int AUX_BYTES = TILE_M * KB * 2;       // 16 KB: stem input / gpool 1x1 input of one group
constexpr int W_STAGE_BYTES = C * KB * 2;        // 8 KB: one tap of one convolution
constexpr int N_STAGES = NG == 1 ? 8 : 16;      // a 3x3 convolution is 9 stages
constexpr int MAX_BLOCKS = 16;
constexpr int MAX_PPT = 8;                       // positions per group
constexpr int MAX_G = 32;                        // gpool channels
constexpr int MAX_P = 32, MAX_H = 64;
constexpr int CAT = C + MAX_P;                    // row stride of the combiner input [f_i | e_i]
constexpr int TMEM_COLS_PER_GROUP = 256;         // X 64 | Y 64 | P 32

struct BlockDesc {
  int gpool;               // 0 res, >0: gpool channels
  const uint8_t* w_pool;   // [g x 64] bf16 SW128 (gpool only)
  const uint8_t* w1;       // 9 x [64 x 64], BN2 folded
  const uint8_t* w2;       // 9 x [64 x 64]
  const float* bn1_s;      // pre-activation scale / shift [64]
  const float* bn1_t;
  const float* b1;         // BN2 shift = bias after conv1 [64]
  const float* pool_s;     // pool_bn scale / shift [64]
  const float* pool_t;
  const float* lin_wT;     // pool_linear weight transposed [2g][64]
  const float* lin_b;      // [64]
};

struct Params {
  const uint8_t* w_stem;   // [64 x 64] (k = tap * 5 + ci, 45 used), stem_bn folded
  const float* b_stem;     // [64]
  const BlockDesc* blocks;
  int n_blocks;
  int width, height, ppt;  // ppt = positions per 128-row group
  int P, H;
  const float* enc_w;      // player_encoder.0.weight [P][3]
  const float* enc_b;      // [P]
  const float* comb_wT;    // combiner.0.weight transposed [C + P][H]
  const float* comb_b;     // [H]
  const float* head_w;     // rows 0-4 policy_head.linear.weight, row 5 value_head.linear.weight: [6][2H]
  const float* head_b;     // [6]
  // per-channel vectors used by every epilogue, contiguous so that they can be staged in shared
  // memory: [b_stem 64] then per block [bn1_s 64][bn1_t 64][b1 64][pool_s 64][pool_t 64]
  const float* vec;
  int vec_floats;          // 64 + 320 * n_blocks
  int vec_in_smem;         // staged (fits next to the operand buffers) or read through L1
};
constexpr int VEC_BLOCK = 320;

struct Smem {
  uint64_t w_full[N_STAGES];
  uint64_t w_empty[N_STAGES];
  uint64_t a_ready[NG];
  uint64_t mma_done[NG];
  uint32_t tmem_base;
};

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float v[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// 16 columns without waiting: several loads can be in flight before one tmem_ld_wait()
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t r[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// named barrier of pipeline `grp` (its GT epilogue threads)
__device__ __forceinline__ void grp_barrier(int grp) { asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(GT) : "memory"); }

// Geometry of the tile row owned by an epilogue thread (padded board layout).
struct RowGeo {
  int grp, lp, cell, x, y;
  bool real;  // the row is a board cell of one of the group's ppt positions (not padding)
};

__device__ __forceinline__ void pack16(const float v[16], uint4& lo, uint4& hi) {
  lo = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  hi = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
}
// 16 channels (16-byte pieces p0, p0 + 1) of buffer row R of a 1024-byte aligned SW128 operand
__device__ __forceinline__ void store_row16(uint8_t* buf, int R, int p0, uint4 lo, uint4 hi) {
  uint8_t* base = buf + (R >> 3) * 1024 + (R & 7) * 128;
  *reinterpret_cast<uint4*>(base + ((p0 ^ (R & 7)) << 4)) = lo;
  *reinterpret_cast<uint4*>(base + (((p0 + 1) ^ (R & 7)) << 4)) = hi;
}
// Descriptor of a [128 x 64] operand that starts `row` rows into a 1024-byte aligned SW128 buffer: the
// hardware applies the 128B swizzle to the shared-memory address, so a start address shifted by whole rows
// reads rows row .. row+127 as they were written (base-offset field 0; checked on B200 by the parity tests).
__device__ __forceinline__ uint64_t desc_rows(uint32_t buf_addr, int row) {
  return umma_desc_sw128(buf_addr + (uint32_t)row * 128u);
}

// out[p][j] = act(bias[j] + sum_k wT[k][j] * in[p][k]) for a handful of rows p (J <= 64 outputs, K <= 96,
// in / out / part in shared memory): the K range is split over GT / J thread groups; a thread
// first requests all of its <= 24 weights (one L2 round trip instead of one per k), then accumulates DENSE_RB
// rows at a time from float4 reads of the inputs; partial sums meet in `part` ([KS][DENSE_RB][J]).
constexpr int DENSE_MAX_KC = 24;  // K <= 96 over GT / J >= 4 thread groups
constexpr int DENSE_RB = 4;       // rows per pass (7x7: 2 positions x 2 players = one pass)
__device__ __forceinline__ void dense_small(const float* __restrict__ wT, const float* __restrict__ bias, int K, int J,
                                            const float* in, int in_stride, int n_rows, float* out, int out_stride,
                                            bool relu, float* part, int tid, int grp) {
  __builtin_assume(__isShared(in));
  __builtin_assume(__isShared(out));
  __builtin_assume(__isShared(part));
  const int KS = GT / J;
  const int Kc = (((K + KS - 1) / KS) + 3) & ~3;  // multiple of 4, <= DENSE_MAX_KC
  const int j = tid % J, kq = tid / J;
  const int k0 = kq * Kc, k1 = min(K, k0 + Kc);
  float w[DENSE_MAX_KC];
#pragma unroll
  for (int i = 0; i < DENSE_MAX_KC; ++i) w[i] = (kq < KS && k0 + i < k1) ? __ldg(wT + (k0 + i) * J + j) : 0.0f;
  for (int p0 = 0; p0 < n_rows; p0 += DENSE_RB) {
    const int np = min(DENSE_RB, n_rows - p0);
    if (kq < KS) {
      float acc[DENSE_RB];
#pragma unroll
      for (int q = 0; q < DENSE_RB; ++q) acc[q] = 0.0f;
      const float* src = in + p0 * in_stride + k0;
#pragma unroll
      for (int g = 0; g < DENSE_MAX_KC / 4; ++g) {
        if (k0 + 4 * g < k1) {
#pragma unroll
          for (int q = 0; q < DENSE_RB; ++q) {
            const float4 x = *reinterpret_cast<const float4*>(src + (q < np ? q : 0) * in_stride + 4 * g);
            // weights past k1 are zero; the matching inputs are zero-filled or finite
            acc[q] += w[4 * g] * x.x + w[4 * g + 1] * x.y + w[4 * g + 2] * x.z + w[4 * g + 3] * x.w;
          }
        }
      }
#pragma unroll
      for (int q = 0; q < DENSE_RB; ++q) part[(kq * DENSE_RB + q) * J + j] = acc[q];
    }
    grp_barrier(grp);
    for (int i = tid; i < np * J; i += GT) {
      const int q = i / J, jj = i - q * J;
      float a = __ldg(bias + jj);
      for (int sp = 0; sp < KS; ++sp) a += part[(sp * DENSE_RB + q) * J + jj];
      out[(p0 + q) * out_stride + jj] = relu ? fmaxf(a, 0.0f) : a;
    }
    grp_barrier(grp);
  }
}

__global__ void __launch_bounds__(THREADS, CTAS_PER_SM)
cnn_forward_kernel(const EvalRow* __restrict__ rows, const uint32_t* __restrict__ n_rows_ptr, int n_rows_arg,
                   const ar_game_pod* __restrict__ games, const uint16_t* __restrict__ maze_tab, Params pr,
                   float* __restrict__ out, int* __restrict__ error_flag) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* a_buf = smem;                                    // NG x 20 KB activations (with guard rows)
  uint8_t* a_aux = a_buf + NG * A_BYTES;                    // NG x 16 KB stem input / gpool 1x1 input; between its
                                                            // uses: pool rows [128][32] f32, head buffers, partials
  uint8_t* ws = a_aux + NG * AUX_BYTES;                     // weight ring
  float* pool_cat = reinterpret_cast<float*>(ws + N_STAGES * W_STAGE_BYTES);  // [NG * MAX_PPT][2 * MAX_G]
  float* pool_out = pool_cat + NG * MAX_PPT * 2 * MAX_G;    // [NG * MAX_PPT][64]
  Smem* sh = reinterpret_cast<Smem*>(pool_out + NG * MAX_PPT * C);
  static_assert(sizeof(Smem) <= 512, "barrier block");
  float* vec_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sh) + 512);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_rows = n_rows_ptr ? (int)*n_rows_ptr : n_rows_arg;
  const int ppt = pr.ppt;
  const int n_tiles = (n_rows + ppt - 1) / ppt;  // one tile = ppt positions = one pass of one pipeline
  if ((int)blockIdx.x * NG >= n_tiles) return;
  const int W = pr.width, Hh = pr.height, PW = W + 1, Sp = PW * (Hh + 1);
  // Round i gives pipeline g of this CTA the tile (i * gridDim.x + blockIdx.x) * NG + g.  The producer and
  // the MMA issuer walk the same job order: for each round, for each job (stem, then per block conv1
  // [+ pool conv] and conv2), for each pipeline that has a tile in this round.
  const int tile_stride = (int)gridDim.x * NG, tile0 = (int)blockIdx.x * NG;

  if (tid == 0) {
    for (int s = 0; s < N_STAGES; ++s) {
      mbar_init(&sh->w_full[s], 1);
      mbar_init(&sh->w_empty[s], 1);
    }
    for (int g = 0; g < NG; ++g) {
      mbar_init(&sh->a_ready[g], GT / 32);
      mbar_init(&sh->mma_done[g], 1);
    }
    fence_barrier_init();
  }
  if (pr.vec_in_smem)
    for (int i = tid; i < pr.vec_floats; i += THREADS) vec_s[i] = __ldg(pr.vec + i);
  const float* vec = pr.vec_in_smem ? vec_s : pr.vec;
  // padding and guard rows are never written: zero the operand buffers once
  for (int i = tid; i < NG * (A_BYTES + AUX_BYTES) / 16; i += THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (warp == EPI_THREADS / 32 + 1) tmem_alloc(&sh->tmem_base, NG * TMEM_COLS_PER_GROUP);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh->tmem_base;
  constexpr uint32_t TX = 0, TY = 64, TP = 128;  // column offsets inside a pipeline's accumulator set

  if (warp == EPI_THREADS / 32) {
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t it = 0;
      auto push = [&](const uint8_t* src, uint32_t bytes) {
        int s = it % N_STAGES;
        uint32_t ph = (it / N_STAGES) & 1;
        mbar_wait(&sh->w_empty[s], ph ^ 1);
        mbar_expect_tx(&sh->w_full[s], bytes);
        tma_bulk_g2s(ws + s * W_STAGE_BYTES, src, bytes, &sh->w_full[s]);
        ++it;
      };
      for (int t0 = tile0; t0 < n_tiles; t0 += tile_stride) {
        const int ng = min(NG, n_tiles - t0);
        for (int g = 0; g < ng; ++g) push(pr.w_stem, W_STAGE_BYTES);
        for (int b = 0; b < pr.n_blocks; ++b) {
          const BlockDesc& bd = pr.blocks[b];
          for (int g = 0; g < ng; ++g) {
            if (bd.gpool) push(bd.w_pool, (uint32_t)bd.gpool * KB * 2);
            for (int tap = 0; tap < 9; ++tap) push(bd.w1 + (size_t)tap * W_STAGE_BYTES, W_STAGE_BYTES);
          }
          for (int g = 0; g < ng; ++g)
            for (int tap = 0; tap < 9; ++tap) push(bd.w2 + (size_t)tap * W_STAGE_BYTES, W_STAGE_BYTES);
        }
      }
    }
  } else if (warp == EPI_THREADS / 32 + 1) {
    // ================= MMA issuer =================
    // The whole warp walks the job list converged (every operand stays warp-uniform, so the descriptors live
    // in uniform registers); the tcgen05 instructions themselves are predicated on lane 0.
    {
      const uint32_t issue = lane == 0 ? 1u : 0u;
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);  // provably warp-uniform
      uint32_t it = 0, a_phase[NG];
      for (int g = 0; g < NG; ++g) a_phase[g] = 0;
      const uint32_t idesc64 = umma_idesc(TILE_M, 64);
      const uint64_t w_desc0 = umma_desc_sw128(smem_u32(ws));
      // one weight stage against pipeline g; tap < 0: the unshifted auxiliary operand
      auto chain = [&](int g, uint32_t d_col, int tap, uint32_t idesc, bool fresh) {
        const int s = __shfl_sync(0xffffffffu, it % N_STAGES, 0);
        const uint32_t ph = (it / N_STAGES) & 1;
        mbar_wait_warp(&sh->w_full[s], ph);
        const uint64_t db = w_desc0 + (uint64_t)(s * (W_STAGE_BYTES >> 4));
        uint64_t da;
        if (tap < 0) {
          da = umma_desc_sw128(smem_u32(a_aux) + g * AUX_BYTES);
        } else {
          const int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
          da = desc_rows(smem_u32(a_buf) + g * A_BYTES, GUARD + dy * PW + dx);
        }
        const uint32_t d = tmem_u + g * TMEM_COLS_PER_GROUP + d_col;
        umma_bf16_pred(d, da, db, idesc, fresh ? 0u : 1u, issue);
#pragma unroll
        for (int k = 1; k < KB / 16; ++k) umma_bf16_pred(d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, 1u, issue);
        umma_commit_pred(&sh->w_empty[s], issue);
        ++it;
      };
      auto wait_a = [&](int g) { mbar_wait_warp(&sh->a_ready[g], a_phase[g]); a_phase[g] ^= 1; tc_fence_after(); };
      for (int t0 = tile0; t0 < n_tiles; t0 += tile_stride) {
        const int ng = min(NG, n_tiles - t0);
        for (int g = 0; g < ng; ++g) {
          wait_a(g);
          chain(g, TX, -1, idesc64, true);  // stem
          umma_commit_pred(&sh->mma_done[g], issue);
        }
        for (int b = 0; b < pr.n_blocks; ++b) {
          const int gp = pr.blocks[b].gpool;
          for (int g = 0; g < ng; ++g) {
            wait_a(g);
            if (gp) chain(g, TP, -1, umma_idesc(TILE_M, gp), true);
#pragma unroll 1
            for (int tap = 0; tap < 9; ++tap) chain(g, TY, tap, idesc64, tap == 0);
            umma_commit_pred(&sh->mma_done[g], issue);
          }
          for (int g = 0; g < ng; ++g) {
            wait_a(g);
#pragma unroll 1
            for (int tap = 0; tap < 9; ++tap) chain(g, TX, tap, idesc64, false);  // += residual
            umma_commit_pred(&sh->mma_done[g], issue);
          }
        }
      }
      __syncwarp();
    }
  } else {
    // ==== encode + epilogues: pipeline grp; threads r and r + 128 own tile row r = TMEM lane r, 32 channels each ====
    const int grp = tid / GT, q = tid - grp * GT, r = q & (TILE_M - 1), half = q >> 7;
    RowGeo g;
    g.lp = r / Sp;
    const int rem = r - g.lp * Sp;
    g.y = rem / PW;
    g.x = rem - g.y * PW;
    g.cell = g.y * W + g.x;
    g.real = g.lp < ppt && g.x < W && g.y < Hh;
    uint8_t* my_buf = a_buf + grp * A_BYTES;
    uint8_t* my_aux = a_aux + grp * AUX_BYTES;
    float* scratch = reinterpret_cast<float*>(my_aux);  // pool rows [128][32] f32 between the operand's uses
    float* pcat = pool_cat + grp * MAX_PPT * 2 * MAX_G;
    float* pout = pool_out + grp * MAX_PPT * C;
    const uint32_t tbase = tmem + grp * TMEM_COLS_PER_GROUP + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t done_phase = 0;
    auto wait_mma = [&]() { mbar_wait(&sh->mma_done[grp], done_phase); done_phase ^= 1; tc_fence_after(); };
    // every thread orders its generic-proxy stores before the async proxy, one lane per warp arrives
    auto signal_a = [&]() {
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->a_ready[grp]);
    };
    float* catb = scratch;                       // [MAX_PPT][2][CAT]   (the operand buffer is free after the trunk)
    float* hid = catb + MAX_PPT * 2 * CAT;       // [MAX_PPT][2][MAX_H]
    float* head_part = hid + MAX_PPT * 2 * MAX_H;  // dense_small partial sums (<= GT * DENSE_RB floats)
    static_assert((MAX_PPT * 2 * (CAT + MAX_H) + GT * DENSE_RB) * 4 <= AUX_BYTES, "head buffers fit the operand buffer");

    for (int t = tile0 + grp; t < n_tiles; t += tile_stride) {
      const int pidx = t * ppt + g.lp;
      const bool live = g.real && pidx < n_rows;
      RowView v;
      const uint16_t* mt = nullptr;
      if (live) {
        const EvalRow er = rows[pidx];
        v = row_view(er, games);
        mt = maze_tab + (size_t)er.game_idx * MAZE_TAB_STRIDE;
      }
      // ---- stem input: im2col of the 5-channel board (4 maze directions from the per-game bf16
      //      table + cheese), k = tap * 5 + ci
      {
        uint16_t e[64];
#pragma unroll
        for (int k = 0; k < 64; ++k) e[k] = 0;
        if (live) {
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3 - 1, dx = tap % 3 - 1;
            const int xn = g.x + dx, yn = g.y + dy;
            if (xn < 0 || xn >= W || yn < 0 || yn >= Hh) continue;
            const int nc = yn * W + xn;
            const uint2 m4 = __ldg(reinterpret_cast<const uint2*>(mt + nc * 4));
            e[tap * 5 + 0] = (uint16_t)(m4.x & 0xffffu);
            e[tap * 5 + 1] = (uint16_t)(m4.x >> 16);
            e[tap * 5 + 2] = (uint16_t)(m4.y & 0xffffu);
            e[tap * 5 + 3] = (uint16_t)(m4.y >> 16);
            e[tap * 5 + 4] = ((v.cheese >> nc) & 1ull) ? BF16_ONE : (uint16_t)0;
          }
        }
#pragma unroll
        for (int p = 0; p < 8; ++p) {
          uint4 pk = make_uint4((uint32_t)e[8 * p] | ((uint32_t)e[8 * p + 1] << 16),
                                (uint32_t)e[8 * p + 2] | ((uint32_t)e[8 * p + 3] << 16),
                                (uint32_t)e[8 * p + 4] | ((uint32_t)e[8 * p + 5] << 16),
                                (uint32_t)e[8 * p + 6] | ((uint32_t)e[8 * p + 7] << 16));
          if ((p >> 2) == half) *reinterpret_cast<uint4*>(my_aux + sw128_offset(r, 8 * p)) = pk;
        }
      }
      signal_a();

      // ---- trunk.  `stage` = -1: stem result; otherwise the result of block `stage`'s conv2.
      for (int stage = -1; stage < pr.n_blocks; ++stage) {
        if (stage >= 0) {
          // conv1 (+ pool conv) of block `stage` finished: y = ReLU(Y + b1) -> activation rows; pool branch
          const BlockDesc& bd = pr.blocks[stage];
          wait_mma();
          uint32_t acc[C / 2];  // this thread's 32 channels: both loads in flight, one wait
#pragma unroll
          for (int cc = 0; cc < C / 2; cc += 16) tmem_ld16_nowait(tbase + TY + half * 32 + cc, acc + cc);
          tmem_ld_wait();
#pragma unroll
          for (int cc = 0; cc < C / 2; cc += 16) {
            const int c0 = half * 32 + cc;
            float y[16];
#pragma unroll
            for (int j = 0; j < 16; ++j)
              y[j] = fmaxf(__uint_as_float(acc[cc + j]) + vec[64 + stage * VEC_BLOCK + 128 + c0 + j], 0.0f);
            uint4 lo, hi;
            pack16(y, lo, hi);
            if (g.real) store_row16(my_buf, GUARD + r, c0 >> 3, lo, hi);
          }
          if (bd.gpool) {
            const int G = bd.gpool;
            for (int c0 = half * 16; c0 < G; c0 += 32) {
              float p[16];
              tmem_ld16(tbase + TP + c0, p);
#pragma unroll
              for (int j = 0; j < 16; j += 4)  // 16-byte chunks XOR-swizzled by row: conflict-free
                *reinterpret_cast<float4*>(scratch + r * MAX_G + ((((c0 + j) >> 2) ^ (r & 7)) << 2)) =
                    make_float4(p[j], p[j + 1], p[j + 2], p[j + 3]);
            }
            grp_barrier(grp);
            for (int i = q; i < ppt * G; i += GT) {  // mean and max over the board, blocks.py:68-70
              const int lq = i / G, ch = i - lq * G;
              float s = 0.0f, m = -INFINITY;
              for (int yy = 0; yy < Hh; ++yy)
                for (int xx = 0; xx < W; ++xx) {
                  const int rr = lq * Sp + yy * PW + xx;
                  float x = scratch[rr * MAX_G + ((((ch >> 2) ^ (rr & 7)) << 2) | (ch & 3))];
                  s += x;
                  m = fmaxf(m, x);
                }
              pcat[lq * 2 * MAX_G + ch] = s / (float)(W * Hh);
              pcat[lq * 2 * MAX_G + G + ch] = m;
            }
            grp_barrier(grp);
            // pool_linear, blocks.py:72 (the pool rows are dead: partial sums reuse their space)
            dense_small(bd.lin_wT, bd.lin_b, 2 * G, C, pcat, 2 * MAX_G, ppt, pout, C, false, scratch, q, grp);
          }
          signal_a();
        }
        // residual stream X is complete for this stage
        wait_mma();
        const bool is_stem = stage < 0;
        const bool add_pool = !is_stem && pr.blocks[stage].gpool != 0;
        const bool last = stage + 1 == pr.n_blocks;
        const BlockDesc* nb = last ? nullptr : &pr.blocks[stage + 1];
        const float* nv = vec + 64 + (stage + 1) * VEC_BLOCK;  // the next block's pre-activation vectors
        uint32_t xacc[C / 2];
#pragma unroll
        for (int cc = 0; cc < C / 2; cc += 16) tmem_ld16_nowait(tbase + TX + half * 32 + cc, xacc + cc);
        tmem_ld_wait();
#pragma unroll
        for (int cc = 0; cc < C / 2; cc += 16) {
          const int c0 = half * 32 + cc;
          float x[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) x[j] = __uint_as_float(xacc[cc + j]);
          if (is_stem) {
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = fmaxf(x[j] + vec[c0 + j], 0.0f);
          }
          if (add_pool && g.real) {
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] += pout[g.lp * C + c0 + j];
          }
          if ((is_stem || add_pool) && !last) tmem_st16(tbase + TX + c0, x);
          if (!last) {
            float a[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) a[j] = fmaxf(x[j] * nv[c0 + j] + nv[64 + c0 + j], 0.0f);
            uint4 lo, hi;
            pack16(a, lo, hi);
            if (g.real) store_row16(my_buf, GUARD + r, c0 >> 3, lo, hi);
            if (nb->gpool) {
#pragma unroll
              for (int j = 0; j < 16; ++j) a[j] = fmaxf(x[j] * nv[192 + c0 + j] + nv[256 + c0 + j], 0.0f);
              pack16(a, lo, hi);
              *reinterpret_cast<uint4*>(my_aux + sw128_offset(r, c0)) = lo;
              *reinterpret_cast<uint4*>(my_aux + sw128_offset(r, c0 + 8)) = hi;
            }
          } else if (live) {
            // trunk finished: features at the players' cells (mask-multiply-sum, model.py:187-190)
            if (g.cell == v.p1)
#pragma unroll
              for (int j = 0; j < 16; ++j) catb[(g.lp * 2 + 0) * CAT + c0 + j] = x[j];
            if (g.cell == v.p2)
#pragma unroll
              for (int j = 0; j < 16; ++j) catb[(g.lp * 2 + 1) * CAT + c0 + j] = x[j];
          }
        }
        if (!last) signal_a();
      }
      tc_fence_before();

      // ---- DeepSet heads on CUDA cores (model.py:192-213); slot sl -> position t * ppt + sl
      const int P = pr.P, H = pr.H;
      const int n_pos = min(ppt, n_rows - t * ppt);
      for (int i = q; i < n_pos * 2 * P; i += GT) {
        const int j = i % P, pl = (i / P) & 1, sl = i / (2 * P);
        RowView q = row_view(rows[t * ppt + sl], games);
        const float side0 = (pl ? q.s2 : q.s1) / 10.0f, side1 = (float)(pl ? q.mud2 : q.mud1) / 10.0f, side2 = q.progress;
        float acc = __ldg(pr.enc_b + j) + __ldg(pr.enc_w + j * 3) * side0 + __ldg(pr.enc_w + j * 3 + 1) * side1 +
                    __ldg(pr.enc_w + j * 3 + 2) * side2;
        catb[(sl * 2 + pl) * CAT + C + j] = fmaxf(acc, 0.0f);
      }
      for (int i = q; i < n_pos * 2 * (MAX_P - P); i += GT)  // the buffer held operand bits: finite padding
        catb[(i / (MAX_P - P)) * CAT + C + P + i % (MAX_P - P)] = 0.0f;
      grp_barrier(grp);
      // combiner: h_i = ReLU(Linear(C + P, H)(cat(f_i, e_i)))
      dense_small(pr.comb_wT, pr.comb_b, C + P, H, catb, CAT, n_pos * 2, hid, MAX_H, true, head_part, q, grp);
      // policy / value heads on cat(h_i, h_1 + h_2): one warp per position, lanes over the hidden units
      for (int sl = (q >> 5); sl < n_pos; sl += GT / 32) {
        float z[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) z[i] = 0.0f;
        for (int k = lane; k < H; k += 32) {
          const float h0 = hid[(sl * 2 + 0) * MAX_H + k], h1 = hid[(sl * 2 + 1) * MAX_H + k], hs = h0 + h1;
#pragma unroll
          for (int a = 0; a < 6; ++a) {
            const float wa = __ldg(pr.head_w + a * 2 * H + k), wb = __ldg(pr.head_w + a * 2 * H + H + k);
            z[a] += wa * h0 + wb * hs;
            z[6 + a] += wa * h1 + wb * hs;
          }
        }
#pragma unroll
        for (int i = 0; i < 12; ++i) {
#pragma unroll
          for (int sh_ = 16; sh_ > 0; sh_ >>= 1) z[i] += __shfl_xor_sync(0xffffffffu, z[i], sh_);
        }
        if (lane == 0) {
          float o[12];
#pragma unroll
          for (int pl = 0; pl < 2; ++pl) {
#pragma unroll
            for (int a = 0; a < 6; ++a) z[pl * 6 + a] += __ldg(pr.head_b + a);
            float m = z[pl * 6];
#pragma unroll
            for (int j = 1; j < 5; ++j) m = fmaxf(m, z[pl * 6 + j]);
            float e[5], sum = 0.0f;
#pragma unroll
            for (int j = 0; j < 5; ++j) { e[j] = expf(z[pl * 6 + j] - m); sum += e[j]; }
#pragma unroll
            for (int j = 0; j < 5; ++j) o[pl * 5 + j] = e[j] / sum;
            float x = z[pl * 6 + 5];
            o[10 + pl] = x > 20.0f ? x : log1pf(expf(x));
          }
          bool ok = true;
#pragma unroll
          for (int j = 0; j < 12; ++j) ok = ok && isfinite(o[j]);
          if (!ok) atomicCAS(error_flag, 0, (int)AR_ERR_NONFINITE);
          float4* dst = reinterpret_cast<float4*>(out + (size_t)(t * ppt + sl) * 12);
          dst[0] = make_float4(o[0], o[1], o[2], o[3]);
          dst[1] = make_float4(o[4], o[5], o[6], o[7]);
          dst[2] = make_float4(o[8], o[9], o[10], o[11]);
        }
      }
      grp_barrier(grp);  // the buffers are reused by the next tile
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == EPI_THREADS / 32 + 1) tmem_dealloc(tmem, NG * TMEM_COLS_PER_GROUP);
}

// ---------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------
struct Model : LeafEvaluator {
  uint8_t* d_blob = nullptr;
  BlockDesc* d_blocks = nullptr;
  Params pr{};
  int n_sms = 148;
  size_t smem_bytes = 0;

  ~Model() override { cudaFree(d_blob); cudaFree(d_blocks); }

  static const float* tensor(const ar_tensor_desc* t, int n, const std::string& name, int64_t numel, std::string& err) {
    const ar_tensor_desc* d = find_tensor(t, n, name.c_str());
    if (!d) { err = "missing tensor " + name; return nullptr; }
    int64_t ne = 1;
    for (int i = 0; i < d->ndim; ++i) ne *= d->shape[i];
    if (numel >= 0 && ne != numel) {
      err = "tensor " + name + " has " + std::to_string(ne) + " elements, expected " + std::to_string(numel);
      return nullptr;
    }
    return d->data;
  }
  // eval-mode BatchNorm2d as y = s * x + t
  static bool bn_affine(const ar_tensor_desc* t, int n, const std::string& bn, int ch, std::vector<float>& s,
                        std::vector<float>& sh, std::string& err) {
    const float* g = tensor(t, n, bn + ".weight", ch, err);
    const float* b = tensor(t, n, bn + ".bias", ch, err);
    const float* mu = tensor(t, n, bn + ".running_mean", ch, err);
    const float* var = tensor(t, n, bn + ".running_var", ch, err);
    if (!g || !b || !mu || !var) return false;
    s.resize(ch); sh.resize(ch);
    for (int c = 0; c < ch; ++c) {
      s[c] = g[c] / sqrtf(var[c] + 1e-5f);
      sh[c] = b[c] - mu[c] * s[c];
    }
    return true;
  }
  // Conv2d weight [64][64][3][3] (optionally scaled per output channel) -> 9 tap images [64 x 64]
  static void conv_taps(const float* w, const float* out_scale, std::vector<uint8_t>& blob) {
    for (int tap = 0; tap < 9; ++tap) {
      std::vector<float> m((size_t)C * C);
      for (int co = 0; co < C; ++co)
        for (int ci = 0; ci < C; ++ci)
          m[(size_t)co * C + ci] = w[((size_t)co * C + ci) * 9 + tap] * (out_scale ? out_scale[co] : 1.0f);
      std::vector<uint8_t> img = swizzled_image(m, C, C, C, 1);
      blob.insert(blob.end(), img.begin(), img.end());
    }
  }

  // A trunk narrower than the kernel's 64 channels runs zero-padded: every tensor with a channel axis is widened
  // with zeros (BatchNorm: weight 0, bias 0, mean 0, variance 1 -> scale 0, shift 0), so the extra channels are exactly
  // 0 in the stem, stay 0 through every pre-activation block (0 + 0 on the residual stream) and meet zero columns
  // in pool_conv and in the combiner.  Same outputs as the narrow network; the 64-channel kernel is unchanged.
  struct PaddedStateDict {
    std::vector<std::vector<float>> store;
    std::vector<ar_tensor_desc> descs;
  };
  static void pad_channels(const ar_tensor_desc* t, int n, int cr, PaddedStateDict& out) {
    auto ends = [](const std::string& s, const char* suf) {
      const size_t m = strlen(suf);
      return s.size() >= m && s.compare(s.size() - m, m, suf) == 0;
    };
    for (int i = 0; i < n; ++i) {
      const ar_tensor_desc& d = t[i];
      const std::string name = d.name ? d.name : "";
      bool pad0 = false, pad1 = false, comb = false;
      float fill = 0.0f;
      if (name == "stem.weight") pad0 = true;
      else if (ends(name, ".conv1.weight") || ends(name, ".conv2.weight")) pad0 = pad1 = true;
      else if (ends(name, ".pool_conv.weight")) pad1 = true;
      else if (ends(name, ".pool_linear.weight") || ends(name, ".pool_linear.bias")) pad0 = true;
      else if (name == "combiner.0.weight") comb = true;
      else if (d.ndim == 1 && d.shape[0] == cr &&
               (name.rfind("stem_bn.", 0) == 0 || name.find(".bn1.") != std::string::npos ||
                name.find(".bn2.") != std::string::npos || name.find(".pool_bn.") != std::string::npos)) {
        pad0 = true;
        fill = ends(name, ".running_var") ? 1.0f : 0.0f;
      }
      if ((pad0 && d.shape[0] != cr) || (pad1 && (d.ndim < 2 || d.shape[1] != cr)) ||
          (comb && (d.ndim != 2 || d.shape[1] <= cr))) {
        pad0 = pad1 = comb = false;  // not the shape this trunk width implies: the loader reports it
      }
      if (!pad0 && !pad1 && !comb) { out.descs.push_back(d); continue; }
      ar_tensor_desc nd = d;
      const int64_t s0 = d.shape[0], s1 = d.ndim >= 2 ? d.shape[1] : 1;
      int64_t inner = 1;
      for (int k = 2; k < d.ndim; ++k) inner *= d.shape[k];
      int64_t n0 = pad0 ? C : s0, n1 = pad1 ? C : s1;
      if (comb) n1 = C + (s1 - cr);
      out.store.emplace_back((size_t)(n0 * n1 * inner), fill);
      float* dst = out.store.back().data();
      for (int64_t a = 0; a < s0; ++a)
        for (int64_t b = 0; b < s1; ++b) {
          const int64_t nb = (comb && b >= cr) ? b - cr + C : b;
          memcpy(dst + (a * n1 + nb) * inner, d.data + (a * s1 + b) * inner, (size_t)inner * sizeof(float));
        }
      nd.data = dst;
      nd.shape[0] = n0;
      if (d.ndim >= 2) nd.shape[1] = n1;
      out.descs.push_back(nd);
    }
  }

  int load(const ar_tensor_desc* t_in, int n_in, int width, int height, std::string& err) override {
    const ar_tensor_desc* t = t_in;
    int n = n_in;
    PaddedStateDict padded;
    {
      const ar_tensor_desc* st0 = find_tensor(t, n, "stem.weight");
      if (st0 && st0->ndim == 4 && st0->shape[0] >= 1 && st0->shape[0] < C) {
        padded.store.reserve((size_t)n);  // the descriptors point into the vectors: no reallocation of the outer one
        pad_channels(t, n, (int)st0->shape[0], padded);
        t = padded.descs.data();
        n = (int)padded.descs.size();
      }
    }
    const int S = width * height;
    if (find_tensor(t, n, "value_head.mlp.0.weight")) { err = "the `pooled` value head has no CUDA evaluator in this build"; return AR_ERR_UNSUPPORTED; }
    const ar_tensor_desc* st = find_tensor(t, n, "stem.weight");
    if (!st || st->ndim != 4) { err = "stem.weight missing (not a PyRatCNN state_dict)"; return AR_ERR_INVALID_ARG; }
    if (st->shape[0] != C || st->shape[1] != 5 || st->shape[2] != 3 || st->shape[3] != 3) {
      err = "the fused CNN kernel needs a 3x3 stem with 5 input and 64 output channels";
      return AR_ERR_UNSUPPORTED;
    }
    int ppt = TILE_M / ((width + 1) * (height + 1));  // padded board rows per position
    if (ppt < 1 || width + 2 > GUARD) { err = "board too large for the fused CNN kernel"; return AR_ERR_UNSUPPORTED; }
    if (ppt > MAX_PPT) ppt = MAX_PPT;
    const ar_tensor_desc* ew = find_tensor(t, n, "player_encoder.0.weight");
    const ar_tensor_desc* cw = find_tensor(t, n, "combiner.0.weight");
    if (!ew || !cw || ew->ndim != 2 || cw->ndim != 2) { err = "player_encoder / combiner missing"; return AR_ERR_INVALID_ARG; }
    const int P = (int)ew->shape[0], H = (int)cw->shape[0];
    if (P > MAX_P || H > MAX_H || ew->shape[1] != 3 || cw->shape[1] != C + P) { err = "unsupported player_dim / hidden_dim"; return AR_ERR_UNSUPPORTED; }

    // The blob holds bf16 operand images (1024-byte aligned) followed by fp32 vectors.
    std::vector<uint8_t> blob;
    std::vector<float> fl;
    auto fpush = [&](const float* p, size_t cnt) { size_t o = fl.size(); fl.insert(fl.end(), p, p + cnt); return o; };

    std::vector<float> ss, sb;
    if (!bn_affine(t, n, "stem_bn", C, ss, sb, err)) return AR_ERR_INVALID_ARG;
    {
      std::vector<float> m((size_t)C * 64, 0.0f);
      for (int co = 0; co < C; ++co)
        for (int ci = 0; ci < 5; ++ci)
          for (int tap = 0; tap < 9; ++tap) m[(size_t)co * 64 + tap * 5 + ci] = st->data[((size_t)co * 5 + ci) * 9 + tap] * ss[co];
      std::vector<uint8_t> img = swizzled_image(m, C, 64, C, 1);
      blob.insert(blob.end(), img.begin(), img.end());
    }
    const size_t o_bstem = fpush(sb.data(), C);

    struct Off { int gpool; size_t w_pool, w1, w2, bn1_s, bn1_t, b1, pool_s, pool_t, lin_wT, lin_b; };
    std::vector<Off> offs;
    for (int b = 0; b < MAX_BLOCKS + 1; ++b) {
      const std::string pre = "blocks." + std::to_string(b);
      const ar_tensor_desc* c1 = find_tensor(t, n, (pre + ".conv1.weight").c_str());
      if (!c1) break;
      if (b == MAX_BLOCKS) { err = "more than 16 trunk blocks"; return AR_ERR_UNSUPPORTED; }
      if (c1->ndim != 4 || c1->shape[0] != C || c1->shape[1] != C || c1->shape[2] != 3) { err = pre + ".conv1.weight: the fused CNN kernel needs 64 channels, 3x3"; return AR_ERR_UNSUPPORTED; }
      const float* w2 = tensor(t, n, pre + ".conv2.weight", (int64_t)C * C * 9, err);
      if (!w2) return AR_ERR_INVALID_ARG;
      std::vector<float> s1, t1, s2, t2;
      if (!bn_affine(t, n, pre + ".bn1", C, s1, t1, err) || !bn_affine(t, n, pre + ".bn2", C, s2, t2, err)) return AR_ERR_INVALID_ARG;
      Off o{};
      const ar_tensor_desc* pc = find_tensor(t, n, (pre + ".pool_conv.weight").c_str());
      if (pc) {
        const int G = (int)pc->shape[0];
        if ((G != 16 && G != 32) || pc->shape[1] != C) { err = pre + ": gpool_channels must be 16 or 32"; return AR_ERR_UNSUPPORTED; }
        std::vector<float> ps, pt;
        if (!bn_affine(t, n, pre + ".pool_bn", C, ps, pt, err)) return AR_ERR_INVALID_ARG;
        const float* lw = tensor(t, n, pre + ".pool_linear.weight", (int64_t)C * 2 * G, err);
        const float* lb = tensor(t, n, pre + ".pool_linear.bias", C, err);
        if (!lw || !lb) return AR_ERR_INVALID_ARG;
        o.gpool = G;
        o.w_pool = blob.size();
        std::vector<float> m(pc->data, pc->data + (size_t)G * C);
        std::vector<uint8_t> img = swizzled_image(m, G, C, G, 1);
        img.resize(1024 * ((img.size() + 1023) / 1024));
        blob.insert(blob.end(), img.begin(), img.end());
        o.pool_s = fpush(ps.data(), C);
        o.pool_t = fpush(pt.data(), C);
        std::vector<float> wT((size_t)2 * G * C);
        for (int c = 0; c < C; ++c)
          for (int j = 0; j < 2 * G; ++j) wT[(size_t)j * C + c] = lw[(size_t)c * 2 * G + j];
        o.lin_wT = fpush(wT.data(), wT.size());
        o.lin_b = fpush(lb, C);
      }
      o.w1 = blob.size();
      conv_taps(c1->data, s2.data(), blob);
      o.w2 = blob.size();
      conv_taps(w2, nullptr, blob);
      o.bn1_s = fpush(s1.data(), C);
      o.bn1_t = fpush(t1.data(), C);
      o.b1 = fpush(t2.data(), C);
      offs.push_back(o);
    }
    if (offs.empty()) { err = "no trunk blocks found"; return AR_ERR_INVALID_ARG; }

    const float* eb = tensor(t, n, "player_encoder.0.bias", P, err);
    const float* cb = tensor(t, n, "combiner.0.bias", H, err);
    const float* pw = tensor(t, n, "policy_head.linear.weight", (int64_t)5 * 2 * H, err);
    const float* pb = tensor(t, n, "policy_head.linear.bias", 5, err);
    const float* vw = tensor(t, n, "value_head.linear.weight", (int64_t)2 * H, err);
    const float* vb = tensor(t, n, "value_head.linear.bias", 1, err);
    if (!eb || !cb || !pw || !pb || !vw || !vb) return AR_ERR_INVALID_ARG;
    const size_t o_encw = fpush(ew->data, (size_t)P * 3), o_encb = fpush(eb, P);
    std::vector<float> cT((size_t)(C + P) * H);
    for (int j = 0; j < H; ++j)
      for (int k = 0; k < C + P; ++k) cT[(size_t)k * H + j] = cw->data[(size_t)j * (C + P) + k];
    const size_t o_combw = fpush(cT.data(), cT.size()), o_combb = fpush(cb, H);
    std::vector<float> hw((size_t)6 * 2 * H);
    memcpy(hw.data(), pw, (size_t)5 * 2 * H * 4);
    memcpy(hw.data() + (size_t)5 * 2 * H, vw, (size_t)2 * H * 4);
    float hb[6] = {pb[0], pb[1], pb[2], pb[3], pb[4], vb[0]};
    const size_t o_headw = fpush(hw.data(), hw.size()), o_headb = fpush(hb, 6);

    std::vector<float> vecs(64 + (size_t)VEC_BLOCK * offs.size(), 0.0f);
    memcpy(&vecs[0], &fl[o_bstem], 64 * 4);
    for (size_t b = 0; b < offs.size(); ++b) {
      float* d = &vecs[64 + b * VEC_BLOCK];
      memcpy(d, &fl[offs[b].bn1_s], 256); memcpy(d + 64, &fl[offs[b].bn1_t], 256); memcpy(d + 128, &fl[offs[b].b1], 256);
      if (offs[b].gpool) { memcpy(d + 192, &fl[offs[b].pool_s], 256); memcpy(d + 256, &fl[offs[b].pool_t], 256); }
    }
    const size_t o_vec = fpush(vecs.data(), vecs.size());
    const size_t fl_off = blob.size();
    blob.resize(fl_off + fl.size() * 4);
    memcpy(blob.data() + fl_off, fl.data(), fl.size() * 4);
#define CKN(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(_e); return AR_ERR_CUDA; } } while (0)
    CKN(cudaMalloc(&d_blob, blob.size()));
    CKN(cudaMemcpy(d_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    const float* dfl = reinterpret_cast<const float*>(d_blob + fl_off);
    std::vector<BlockDesc> descs(offs.size());
    for (size_t b = 0; b < offs.size(); ++b) {
      const Off& o = offs[b];
      BlockDesc& d = descs[b];
      d.gpool = o.gpool;
      d.w_pool = o.gpool ? d_blob + o.w_pool : nullptr;
      d.w1 = d_blob + o.w1; d.w2 = d_blob + o.w2;
      d.bn1_s = dfl + o.bn1_s; d.bn1_t = dfl + o.bn1_t; d.b1 = dfl + o.b1;
      d.pool_s = dfl + o.pool_s; d.pool_t = dfl + o.pool_t; d.lin_wT = dfl + o.lin_wT; d.lin_b = dfl + o.lin_b;
    }
    CKN(cudaMalloc(&d_blocks, descs.size() * sizeof(BlockDesc)));
    CKN(cudaMemcpy(d_blocks, descs.data(), descs.size() * sizeof(BlockDesc), cudaMemcpyHostToDevice));
    pr.w_stem = d_blob;
    pr.b_stem = dfl + o_bstem;
    pr.vec = dfl + o_vec;
    pr.vec_floats = (int)vecs.size();
    pr.blocks = d_blocks;
    pr.n_blocks = (int)descs.size();
    pr.width = width; pr.height = height; pr.ppt = ppt;
    pr.P = P; pr.H = H;
    pr.enc_w = dfl + o_encw; pr.enc_b = dfl + o_encb;
    pr.comb_wT = dfl + o_combw; pr.comb_b = dfl + o_combb;
    pr.head_w = dfl + o_headw; pr.head_b = dfl + o_headb;
    smem_bytes = (size_t)NG * (A_BYTES + AUX_BYTES) + N_STAGES * W_STAGE_BYTES +
                 (size_t)NG * MAX_PPT * 2 * MAX_G * 4 + (size_t)NG * MAX_PPT * C * 4 + 512 + 1024;
    const size_t smem_limit = (size_t)228 * 1024 / CTAS_PER_SM - 1024;  // per CTA, CTAS_PER_SM resident
    pr.vec_in_smem = smem_bytes + vecs.size() * 4 <= smem_limit ? 1 : 0;
    if (pr.vec_in_smem) smem_bytes += vecs.size() * 4;
    CKN(cudaFuncSetAttribute(cnn_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    int dev = 0;
    CKN(cudaGetDevice(&dev));
    CKN(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
#undef CKN
    return AR_OK;
  }

  cudaError_t forward(const EvalRow* rows, const uint32_t* n_rows_dev, int n_rows_max, const ar_game_pod* games,
                      const uint16_t* maze_tab, float* out, int* error_flag, cudaStream_t stream) const override {
    if (n_rows_max <= 0) return cudaSuccess;
    int tiles = (n_rows_max + pr.ppt - 1) / pr.ppt;
    int ctas = (tiles + NG - 1) / NG;
    int grid = ctas < n_sms * CTAS_PER_SM ? ctas : n_sms * CTAS_PER_SM;
    cnn_forward_kernel<<<grid, THREADS, smem_bytes, stream>>>(rows, n_rows_dev, n_rows_max, games, maze_tab, pr, out,
                                                              error_flag);
    return cudaGetLastError();
  }
};

}  // namespace cnn

LeafEvaluator* make_cnn_evaluator() { return new cnn::Model(); }

}  // namespace ar
