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
// B200 mapping: a 3x3 convolution over a tile of `ppt` positions (ppt * S <= 128 spatial rows, one
// row per TMEM lane) is an implicit GEMM with M = 128, N = 64, K = 9 taps x 64 channels.  The A
// operand is an im2col image in shared memory: 9 K-blocks of [128 x 64] bf16, 128B-swizzled, where
// block `tap` row r' holds the activations of r' + tap offset (zero outside the board).  The
// epilogue thread that owns a row scatters its 128-byte activation vector into the 9 tap blocks.
// The residual stream never leaves TMEM: conv2 accumulates straight onto it (accumulate = 1 from
// the first MMA), conv1 goes to a second accumulator, the gpool 1x1 convolution to a third.
// Weights stream tap by tap (8 KB) through a 4-stage TMA bulk-copy ring.  BatchNorms that follow
// a convolution are folded into its weights; pre-activation BatchNorms are applied in the epilogue.
// The tiny per-player head (3 -> P -> H -> 6) runs on CUDA cores from shared memory.
#include "nn_common.cuh"

namespace ar {
namespace cnn {

constexpr int TILE_M = 128;
constexpr int C = 64;
constexpr int A_BLOCK_BYTES = TILE_M * KB * 2;   // 16 KB: one tap of the im2col image
constexpr int W_STAGE_BYTES = C * KB * 2;        // 8 KB: one tap of one convolution
constexpr int N_STAGES = 4;
constexpr int THREADS = 192;
constexpr int MAX_BLOCKS = 16;
constexpr int MAX_PPT = 8;                       // positions per tile (boards smaller than 4x4 waste rows)
constexpr int MAX_G = 32;                        // gpool channels
constexpr int MAX_P = 32, MAX_H = 64;

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
  int width, height, ppt;
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
  uint64_t a_ready;
  uint64_t mma_done;
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
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// Geometry of the tile row owned by an epilogue thread.
struct RowGeo {
  int lp, cell, x, y;
  bool in_tile;  // row belongs to one of the tile's ppt positions
};

// Scatter 16 channels (pieces p0, p0+1 of the 128-byte activation vector of row r) into the 9 tap
// blocks of the im2col image: block `tap` row r' = r - (dy * w + dx) holds in[r' + (dy, dx)].
__device__ __forceinline__ void scatter_taps(uint8_t* a_taps, const RowGeo& g, int r, int w, int h, int p0,
                                             uint4 lo, uint4 hi) {
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
    const int xo = g.x - dx, yo = g.y - dy;  // the output cell that reads this row through `tap`
    if (xo < 0 || xo >= w || yo < 0 || yo >= h) continue;
    const int ro = r - (dy * w + dx);
    uint8_t* base = a_taps + tap * A_BLOCK_BYTES + (ro >> 3) * 1024 + (ro & 7) * 128;
    *reinterpret_cast<uint4*>(base + (((p0) ^ (ro & 7)) << 4)) = lo;
    *reinterpret_cast<uint4*>(base + (((p0 + 1) ^ (ro & 7)) << 4)) = hi;
  }
}

__device__ __forceinline__ void pack16(const float v[16], uint4& lo, uint4& hi) {
  lo = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  hi = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
}

__global__ void __launch_bounds__(THREADS, 1)
cnn_forward_kernel(const EvalRow* __restrict__ rows, const uint32_t* __restrict__ n_rows_ptr, int n_rows_arg,
                   const ar_game_pod* __restrict__ games, const uint16_t* __restrict__ maze_tab, Params pr,
                   float* __restrict__ out, int* __restrict__ error_flag) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* a_taps = smem;                                   // 9 x 16 KB im2col image
  uint8_t* a_aux = smem + 9 * A_BLOCK_BYTES;                // stem input / gpool 1x1 input, [128 x 64]
  uint8_t* ws = a_aux + A_BLOCK_BYTES;                      // weight ring
  float* scratch = reinterpret_cast<float*>(ws + N_STAGES * W_STAGE_BYTES);  // [128][32] pool rows; head buffers
  float* pool_cat = scratch + TILE_M * MAX_G;               // [MAX_PPT][2 * MAX_G]
  float* pool_out = pool_cat + MAX_PPT * 2 * MAX_G;         // [MAX_PPT][64]
  Smem* sh = reinterpret_cast<Smem*>(pool_out + MAX_PPT * C);
  float* vec_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sh) + 128);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_rows = n_rows_ptr ? (int)*n_rows_ptr : n_rows_arg;
  const int ppt = pr.ppt;
  const int n_tiles = (n_rows + ppt - 1) / ppt;
  if ((int)blockIdx.x >= n_tiles) return;
  const int W = pr.width, Hh = pr.height, S = W * Hh;

  if (tid == 0) {
    for (int s = 0; s < N_STAGES; ++s) {
      mbar_init(&sh->w_full[s], 1);
      mbar_init(&sh->w_empty[s], 1);
    }
    mbar_init(&sh->a_ready, 128);
    mbar_init(&sh->mma_done, 1);
    fence_barrier_init();
  }
  if (pr.vec_in_smem)
    for (int i = tid; i < pr.vec_floats; i += THREADS) vec_s[i] = __ldg(pr.vec + i);
  const float* vec = pr.vec_in_smem ? vec_s : pr.vec;
  // out-of-board taps are never written: zero the im2col image once
  for (int i = tid; i < 10 * A_BLOCK_BYTES / 16; i += THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 5) tmem_alloc(&sh->tmem_base, 256);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh->tmem_base;
  const uint32_t TX = tmem, TY = tmem + 64, TP = tmem + 128;

  if (warp == 4) {
    // ================= TMA producer: stem, then per block [pool conv,] conv1 taps, conv2 taps =====
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
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        push(pr.w_stem, W_STAGE_BYTES);
        for (int b = 0; b < pr.n_blocks; ++b) {
          const BlockDesc& bd = pr.blocks[b];
          if (bd.gpool) push(bd.w_pool, (uint32_t)bd.gpool * KB * 2);
          for (int tap = 0; tap < 9; ++tap) push(bd.w1 + (size_t)tap * W_STAGE_BYTES, W_STAGE_BYTES);
          for (int tap = 0; tap < 9; ++tap) push(bd.w2 + (size_t)tap * W_STAGE_BYTES, W_STAGE_BYTES);
        }
      }
    }
  } else if (warp == 5) {
    // ================= MMA issuer =================
    if (lane == 0) {
      uint32_t it = 0, a_phase = 0;
      const uint32_t idesc64 = umma_idesc(TILE_M, 64);
      auto chain = [&](uint32_t d_tmem, const uint8_t* a_block, uint32_t idesc, bool fresh) {
        int s = it % N_STAGES;
        uint32_t ph = (it / N_STAGES) & 1;
        mbar_wait(&sh->w_full[s], ph);
        tc_fence_after();
        uint64_t da = umma_desc_sw128(smem_u32(a_block));
        uint64_t db = umma_desc_sw128(smem_u32(ws + s * W_STAGE_BYTES));
#pragma unroll
        for (int k = 0; k < KB / 16; ++k)
          umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (fresh && k == 0) ? 0u : 1u);
        umma_commit(&sh->w_empty[s]);
        ++it;
      };
      auto wait_a = [&]() { mbar_wait(&sh->a_ready, a_phase); a_phase ^= 1; tc_fence_after(); };
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        wait_a();
        chain(TX, a_aux, idesc64, true);  // stem
        umma_commit(&sh->mma_done);
        for (int b = 0; b < pr.n_blocks; ++b) {
          const int g = pr.blocks[b].gpool;
          wait_a();
          if (g) chain(TP, a_aux, umma_idesc(TILE_M, g), true);
          for (int tap = 0; tap < 9; ++tap) chain(TY, a_taps + tap * A_BLOCK_BYTES, idesc64, tap == 0);
          umma_commit(&sh->mma_done);
          wait_a();
          for (int tap = 0; tap < 9; ++tap) chain(TX, a_taps + tap * A_BLOCK_BYTES, idesc64, false);  // += residual
          umma_commit(&sh->mma_done);
        }
      }
    }
  } else {
    // ================= encode + epilogues: thread r owns tile row r = TMEM lane r =================
    const int r = tid;
    RowGeo g;
    g.lp = r / S;
    g.cell = r - g.lp * S;
    g.y = g.cell / W;
    g.x = g.cell - g.y * W;
    g.in_tile = g.lp < ppt;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    uint32_t done_phase = 0;
    auto wait_mma = [&]() { mbar_wait(&sh->mma_done, done_phase); done_phase ^= 1; tc_fence_after(); };
    auto signal_a = [&]() { tc_fence_before(); fence_proxy_async(); mbar_arrive(&sh->a_ready); };
    float* feat = scratch;                       // [MAX_PPT][2][64]   (scratch is free after the trunk)
    float* enc = feat + MAX_PPT * 2 * C;         // [MAX_PPT][2][MAX_P]
    float* hid = enc + MAX_PPT * 2 * MAX_P;      // [MAX_PPT][2][MAX_H]
    float* zbuf = hid + MAX_PPT * 2 * MAX_H;     // [MAX_PPT][12]

    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int pidx = t * ppt + g.lp;
      const bool live = g.in_tile && pidx < n_rows;
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
          *reinterpret_cast<uint4*>(a_aux + sw128_offset(r, 8 * p)) = pk;
        }
      }
      signal_a();

      // ---- trunk.  `stage` = -1: stem result; otherwise the result of block `stage`'s conv2.
      for (int stage = -1; stage < pr.n_blocks; ++stage) {
        if (stage >= 0) {
          // conv1 (+ pool conv) of block `stage` finished: y = ReLU(Y + b1) -> im2col; pool branch
          const BlockDesc& bd = pr.blocks[stage];
          wait_mma();
#pragma unroll 1
          for (int c0 = 0; c0 < C; c0 += 16) {
            float y[16];
            tmem_ld16(TY + lane_off + c0, y);
#pragma unroll
            for (int j = 0; j < 16; ++j) y[j] = fmaxf(y[j] + vec[64 + stage * VEC_BLOCK + 128 + c0 + j], 0.0f);
            uint4 lo, hi;
            pack16(y, lo, hi);
            if (g.in_tile) scatter_taps(a_taps, g, r, W, Hh, c0 >> 3, lo, hi);
          }
          if (bd.gpool) {
            const int G = bd.gpool;
            for (int c0 = 0; c0 < G; c0 += 16) {
              float p[16];
              tmem_ld16(TP + lane_off + c0, p);
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<float4*>(scratch + r * MAX_G + c0 + j) = make_float4(p[j], p[j + 1], p[j + 2], p[j + 3]);
            }
            epi_barrier();
            for (int i = tid; i < ppt * G; i += 128) {  // mean and max over the board, blocks.py:68-70
              const int lp = i / G, ch = i - lp * G;
              float s = 0.0f, m = -INFINITY;
              for (int c = 0; c < S; ++c) {
                float x = scratch[(lp * S + c) * MAX_G + ch];
                s += x;
                m = fmaxf(m, x);
              }
              pool_cat[lp * 2 * MAX_G + ch] = s / (float)S;
              pool_cat[lp * 2 * MAX_G + G + ch] = m;
            }
            epi_barrier();
            for (int i = tid; i < ppt * C; i += 128) {  // pool_linear, blocks.py:72
              const int lp = i >> 6, c = i & 63;
              float acc = __ldg(bd.lin_b + c);
              for (int j = 0; j < 2 * G; ++j) acc += __ldg(bd.lin_wT + j * C + c) * pool_cat[lp * 2 * MAX_G + j];
              pool_out[lp * C + c] = acc;
            }
            epi_barrier();
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
#pragma unroll 1
        for (int c0 = 0; c0 < C; c0 += 16) {
          float x[16];
          tmem_ld16(TX + lane_off + c0, x);
          if (is_stem) {
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = fmaxf(x[j] + vec[c0 + j], 0.0f);
          }
          if (add_pool && g.in_tile) {
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] += pool_out[g.lp * C + c0 + j];
          }
          if ((is_stem || add_pool) && !last) tmem_st16(TX + lane_off + c0, x);
          if (!last) {
            float a[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) a[j] = fmaxf(x[j] * nv[c0 + j] + nv[64 + c0 + j], 0.0f);
            uint4 lo, hi;
            pack16(a, lo, hi);
            if (g.in_tile) scatter_taps(a_taps, g, r, W, Hh, c0 >> 3, lo, hi);
            if (nb->gpool) {
#pragma unroll
              for (int j = 0; j < 16; ++j) a[j] = fmaxf(x[j] * nv[192 + c0 + j] + nv[256 + c0 + j], 0.0f);
              pack16(a, lo, hi);
              *reinterpret_cast<uint4*>(a_aux + sw128_offset(r, c0)) = lo;
              *reinterpret_cast<uint4*>(a_aux + sw128_offset(r, c0 + 8)) = hi;
            }
          } else if (live) {
            // trunk finished: features at the players' cells (mask-multiply-sum, model.py:187-190)
            if (g.cell == v.p1)
#pragma unroll
              for (int j = 0; j < 16; ++j) feat[(g.lp * 2 + 0) * C + c0 + j] = x[j];
            if (g.cell == v.p2)
#pragma unroll
              for (int j = 0; j < 16; ++j) feat[(g.lp * 2 + 1) * C + c0 + j] = x[j];
          }
        }
        if (!last) signal_a();
      }
      tc_fence_before();

      // ---- DeepSet heads on CUDA cores (model.py:192-213)
      const int P = pr.P, H = pr.H;
      const int n_pos = min(ppt, n_rows - t * ppt);
      for (int i = tid; i < n_pos * 2 * P; i += 128) {
        const int j = i % P, pl = (i / P) & 1, lp = i / (2 * P);
        RowView q = row_view(rows[t * ppt + lp], games);
        const float side0 = (pl ? q.s2 : q.s1) / 10.0f, side1 = (float)(pl ? q.mud2 : q.mud1) / 10.0f, side2 = q.progress;
        float acc = __ldg(pr.enc_b + j) + __ldg(pr.enc_w + j * 3) * side0 + __ldg(pr.enc_w + j * 3 + 1) * side1 +
                    __ldg(pr.enc_w + j * 3 + 2) * side2;
        enc[(lp * 2 + pl) * MAX_P + j] = fmaxf(acc, 0.0f);
      }
      epi_barrier();
      for (int i = tid; i < n_pos * 2 * H; i += 128) {
        const int j = i % H, pl = (i / H) & 1, lp = i / (2 * H);
        const float* f = feat + (lp * 2 + pl) * C;
        const float* e = enc + (lp * 2 + pl) * MAX_P;
        float acc = __ldg(pr.comb_b + j);
        for (int k = 0; k < C; ++k) acc += __ldg(pr.comb_wT + k * H + j) * f[k];
        for (int k = 0; k < P; ++k) acc += __ldg(pr.comb_wT + (C + k) * H + j) * e[k];
        hid[(lp * 2 + pl) * MAX_H + j] = fmaxf(acc, 0.0f);
      }
      epi_barrier();
      for (int i = tid; i < n_pos * 12; i += 128) {
        const int a = i % 6, pl = (i / 6) & 1, lp = i / 12;
        const float* hi_ = hid + (lp * 2 + pl) * MAX_H;
        const float* ho = hid + (lp * 2 + (pl ^ 1)) * MAX_H;
        const float* wr = pr.head_w + a * 2 * H;
        float acc = __ldg(pr.head_b + a);
        for (int k = 0; k < H; ++k) acc += __ldg(wr + k) * hi_[k] + __ldg(wr + H + k) * (hi_[k] + ho[k]);
        zbuf[lp * 12 + pl * 6 + a] = acc;
      }
      epi_barrier();
      if (tid < n_pos) {
        const float* z = zbuf + tid * 12;
        float o[12];
#pragma unroll
        for (int pl = 0; pl < 2; ++pl) {
          float m = z[pl * 6];
#pragma unroll
          for (int j = 1; j < 5; ++j) m = fmaxf(m, z[pl * 6 + j]);
          float e[5], s = 0.0f;
#pragma unroll
          for (int j = 0; j < 5; ++j) { e[j] = expf(z[pl * 6 + j] - m); s += e[j]; }
#pragma unroll
          for (int j = 0; j < 5; ++j) o[pl * 5 + j] = e[j] / s;
          float x = z[pl * 6 + 5];
          o[10 + pl] = x > 20.0f ? x : log1pf(expf(x));
        }
        bool ok = true;
#pragma unroll
        for (int j = 0; j < 12; ++j) ok = ok && isfinite(o[j]);
        if (!ok) atomicCAS(error_flag, 0, (int)AR_ERR_NONFINITE);
        float4* dst = reinterpret_cast<float4*>(out + (size_t)(t * ppt + tid) * 12);
        dst[0] = make_float4(o[0], o[1], o[2], o[3]);
        dst[1] = make_float4(o[4], o[5], o[6], o[7]);
        dst[2] = make_float4(o[8], o[9], o[10], o[11]);
      }
      epi_barrier();  // scratch is reused by the next tile
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem, 256);
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

  int load(const ar_tensor_desc* t, int n, int width, int height, std::string& err) override {
    const int S = width * height;
    if (find_tensor(t, n, "value_head.mlp.0.weight")) { err = "the `pooled` value head has no CUDA evaluator in this build"; return AR_ERR_UNSUPPORTED; }
    const ar_tensor_desc* st = find_tensor(t, n, "stem.weight");
    if (!st || st->ndim != 4) { err = "stem.weight missing (not a PyRatCNN state_dict)"; return AR_ERR_INVALID_ARG; }
    if (st->shape[0] != C || st->shape[1] != 5 || st->shape[2] != 3 || st->shape[3] != 3) {
      err = "the fused CNN kernel needs a 3x3 stem with 5 input and 64 output channels";
      return AR_ERR_UNSUPPORTED;
    }
    int ppt = TILE_M / S;
    if (ppt < 1) { err = "board too large for the fused CNN kernel"; return AR_ERR_UNSUPPORTED; }
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
    smem_bytes = (size_t)10 * A_BLOCK_BYTES + N_STAGES * W_STAGE_BYTES + (size_t)TILE_M * MAX_G * 4 +
                 (size_t)MAX_PPT * 2 * MAX_G * 4 + (size_t)MAX_PPT * C * 4 + 128 + 1024;
    pr.vec_in_smem = smem_bytes + vecs.size() * 4 <= 227 * 1024 ? 1 : 0;
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
    int grid = tiles < n_sms ? tiles : n_sms;
    cnn_forward_kernel<<<grid, THREADS, smem_bytes, stream>>>(rows, n_rows_dev, n_rows_max, games, maze_tab, pr, out,
                                                              error_flag);
    return cudaGetLastError();
  }
};

}  // namespace cnn

LeafEvaluator* make_cnn_evaluator() { return new cnn::Model(); }

}  // namespace ar
