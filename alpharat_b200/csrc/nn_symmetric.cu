// nn_symmetric.cu — SymmetricMLP (DeepSet) leaf evaluator on tcgen05 tensor cores.
//
// Graph: SymmetricMLP.predict, alpharat/nn/models/symmetric.py:124-229 (eval mode):
//   shared = ReLU(BN(Linear(5S+1, 256)([maze, cheese, progress])))
//   p_i    = ReLU(BN(Linear(S+2, 256)([pos_i, mud_i, score_i])))          same encoder for both players
//   h_i    = trunk(cat(shared, p_i))      trunk = Linear(512,256) BN ReLU Linear(256,256) BN ReLU
//   agg    = h_1 + h_2
//   logits_i = policy_head(cat(h_i, agg)) -> softmax ; value_i = softplus(value_head(cat(h_i, agg)))
//
// B200 mapping: the weight sharing across players is turned into GEMM rows — one 128-row tile holds
// 64 positions x 2 players (row 2p + i), so every layer is one M=128 tcgen05.mma chain and the
// DeepSet sum needs only the neighbouring row (one shuffle between lanes t and t^1):
//   stage A  acc0 = Ws  . shared_raw   (K = ceil((5S+1)/64) blocks)    acc1 = Wp . player_raw (1 block)
//   stage B  acc0 = Wt1 . [shared | p]  (K = 512)
//   stage C  acc0 = Wt2 . t             (K = 256)
//   stage D  acc0 = Wh  . h             (N = 16: rows 0-5 act on h_i, rows 6-11 on agg)
//            cat(h_i, agg).W = Wa.h_i + Wb.(h_i + h_partner), combined in fp32 in the epilogue.
// Observations are encoded straight into the swizzled bf16 A operand (never materialised in HBM);
// weights stream through a 3-stage TMA bulk-copy ring; accumulators live in TMEM (512 columns).
// Numerics as in nn_kernels.cu: bf16 operands, fp32 accumulation, fp32 heads.
#include "nn_common.cuh"

namespace ar {
namespace sym {

constexpr int TILE_M = 128;
constexpr int POS_PER_TILE = 64;
constexpr int A_BLOCK_BYTES = TILE_M * KB * 2;  // 16 KB
constexpr int W_STAGE_BYTES = 256 * KB * 2;     // 32 KB
constexpr int WH_CHUNK_BYTES = 16 * KB * 2;     // 2 KB
constexpr int N_STAGES = 3;
constexpr int EPI_WARPS = 8;                    // two threads per tile row, 128 output columns each
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int THREADS = EPI_THREADS + 64;       // + TMA warp + MMA warp

struct Weights {
  const uint8_t* ws;   // ks_blocks x [256 x 64]
  const uint8_t* wp;   // 1 x [256 x 64]
  const uint8_t* wt1;  // 8 x [256 x 64]
  const uint8_t* wt2;  // 4 x [256 x 64]
  const uint8_t* wh;   // 4 x [16 x 64]
  const float* bs;
  const float* bp;
  const float* bt1;
  const float* bt2;
  const float* bh;     // policy bias[5], value bias
  int ks_blocks;
};

struct Smem {
  uint64_t w_full[N_STAGES];
  uint64_t w_empty[N_STAGES];
  uint64_t a_ready;
  uint64_t mma_done;
  uint32_t tmem_base;
};

// shared_raw = [maze(4S), cheese(S), progress] — symmetric.py:137-147 (assembled in the kernel)
// [pos one-hot(S), mud, score] — symmetric.py:150-165
__device__ __forceinline__ float player_elem(const RowView& v, int k, int player) {
  const int S = v.spatial;
  const int pos = player ? v.p2 : v.p1;
  if (k < S) return k == pos ? 1.0f : 0.0f;
  if (k == S) return (float)(player ? v.mud2 : v.mud1) / 10.0f;
  if (k == S + 1) return (player ? v.s2 : v.s1) / 10.0f;
  return 0.0f;
}

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

// TMEM accumulator (this thread's 128 of the 256 fp32 columns of its lane) -> +bias, ReLU -> bf16 A operand
__device__ __forceinline__ void hidden_epilogue(uint32_t t_addr, const float* __restrict__ bias, uint8_t* dst,
                                                int r, int half) {
#pragma unroll 1
  for (int c64 = half * 128; c64 < half * 128 + 128; c64 += 64) {
    uint32_t acc[64];  // four loads in flight, one wait
#pragma unroll
    for (int cc = 0; cc < 64; cc += 16) tmem_ld16_nowait(t_addr + c64 + cc, acc + cc);
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
      uint4 p0 = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
      uint4 p1 = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]),
                            pack_bf16(v[14], v[15]));
      *reinterpret_cast<uint4*>(dst + kb * A_BLOCK_BYTES + sw128_offset(r, col)) = p0;
      *reinterpret_cast<uint4*>(dst + kb * A_BLOCK_BYTES + sw128_offset(r, col + 8)) = p1;
    }
  }
}

// warps 0-7: encode + epilogues (threads r and r + 128 own tile row r = TMEM lane r, 128 columns each),
// warp 8: TMA producer, warp 9: MMA issuer (converged; tcgen05 instructions predicated on lane 0) and TMEM owner.
__global__ void __launch_bounds__(THREADS, 1)
symmetric_forward_kernel(const EvalRow* __restrict__ rows, const uint32_t* __restrict__ n_rows_ptr, int n_rows_arg,
                         const ar_game_pod* __restrict__ games, const uint16_t* __restrict__ maze_tab, Weights w,
                         float* __restrict__ out, int* __restrict__ error_flag) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* r0 = smem;                           // shared_raw -> shared -> t
  uint8_t* r1 = smem + 4 * A_BLOCK_BYTES;       // player_raw -> p -> h
  uint8_t* ws = smem + 8 * A_BLOCK_BYTES;       // weight stages
  Smem* sh = reinterpret_cast<Smem*>(ws + N_STAGES * W_STAGE_BYTES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_rows = n_rows_ptr ? (int)*n_rows_ptr : n_rows_arg;
  const int n_tiles = (n_rows + POS_PER_TILE - 1) / POS_PER_TILE;
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
  if (warp == EPI_WARPS + 1) tmem_alloc(&sh->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh->tmem_base;
  const int ks = w.ks_blocks;
  const int chunks_per_tile = ks + 1 + 8 + 4 + 4;

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
          uint32_t bytes = W_STAGE_BYTES;
          int cc = c;
          if (cc < ks) src = w.ws + (size_t)cc * W_STAGE_BYTES;
          else if ((cc -= ks) < 1) src = w.wp;
          else if ((cc -= 1) < 8) src = w.wt1 + (size_t)cc * W_STAGE_BYTES;
          else if ((cc -= 8) < 4) src = w.wt2 + (size_t)cc * W_STAGE_BYTES;
          else { cc -= 4; src = w.wh + (size_t)cc * WH_CHUNK_BYTES; bytes = WH_CHUNK_BYTES; }
          mbar_expect_tx(&sh->w_full[s], bytes);
          tma_bulk_g2s(ws + s * W_STAGE_BYTES, src, bytes, &sh->w_full[s]);
        }
      }
    }
  } else if (warp == EPI_WARPS + 1) {
    // ================= MMA issuer (whole warp converged) =================
    {
      const uint32_t issue = lane == 0 ? 1u : 0u;
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
      uint32_t it = 0, a_phase = 0;
      const uint32_t idesc256 = umma_idesc(TILE_M, 256), idesc16 = umma_idesc(TILE_M, 16);
      const uint64_t w_desc0 = umma_desc_sw128(smem_u32(ws));
      auto chain = [&](uint32_t d_tmem, const uint8_t* a_block, uint32_t idesc, bool first) {
        const int s = __shfl_sync(0xffffffffu, it % N_STAGES, 0);
        const uint32_t ph = (it / N_STAGES) & 1;
        mbar_wait_warp(&sh->w_full[s], ph);
        const uint64_t da = umma_desc_sw128(smem_u32(a_block));
        const uint64_t db = w_desc0 + (uint64_t)(s * (W_STAGE_BYTES >> 4));
        umma_bf16_pred(d_tmem, da, db, idesc, first ? 0u : 1u, issue);
#pragma unroll
        for (int k = 1; k < KB / 16; ++k) umma_bf16_pred(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, 1u, issue);
        umma_commit_pred(&sh->w_empty[s], issue);
        ++it;
      };
      auto wait_a = [&]() { mbar_wait_warp(&sh->a_ready, a_phase); a_phase ^= 1; tc_fence_after(); };
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        // stage A: the two encoders into the two accumulators
        wait_a();
        for (int kb = 0; kb < ks; ++kb) chain(tmem_u, r0 + kb * A_BLOCK_BYTES, idesc256, kb == 0);
        chain(tmem_u + 256, r1, idesc256, true);
        umma_commit_pred(&sh->mma_done, issue);
        // stage B: trunk layer 1 over cat(shared, p)
        wait_a();
        for (int kb = 0; kb < 8; ++kb)
          chain(tmem_u, (kb < 4 ? r0 + kb * A_BLOCK_BYTES : r1 + (kb - 4) * A_BLOCK_BYTES), idesc256, kb == 0);
        umma_commit_pred(&sh->mma_done, issue);
        // stage C: trunk layer 2
        wait_a();
        for (int kb = 0; kb < 4; ++kb) chain(tmem_u, r0 + kb * A_BLOCK_BYTES, idesc256, kb == 0);
        umma_commit_pred(&sh->mma_done, issue);
        // stage D: heads
        wait_a();
        for (int kb = 0; kb < 4; ++kb) chain(tmem_u, r1 + kb * A_BLOCK_BYTES, idesc16, kb == 0);
        umma_commit_pred(&sh->mma_done, issue);
      }
      __syncwarp();
    }
  } else {
    // ================= encode + epilogues =================
    const int r = tid & (TILE_M - 1), half = tid >> 7;
    const int player = r & 1;
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
      const int pidx = t * POS_PER_TILE + (r >> 1);
      const bool live = pidx < n_rows;
      {
        // shared_raw = [maze (copied from the per-game table), cheese, progress]; player_raw =
        // [position one-hot, mud, score]: fill with 16-byte stores, then place the few non-zeros
        RowView v;
        const uint4* mt = nullptr;
        int n_maze_pieces = 0;
        if (live) {
          const EvalRow er = rows[pidx];
          v = row_view(er, games);
          mt = reinterpret_cast<const uint4*>(maze_tab + (size_t)er.game_idx * MAZE_TAB_STRIDE);
          n_maze_pieces = (4 * v.spatial + 7) >> 3;
        }
        for (int p = half * ks * 4; p < (half + 1) * ks * 4; ++p) {  // each thread of the pair fills half of the row
          uint4 pk = make_uint4(0, 0, 0, 0);
          if (p < n_maze_pieces) pk = __ldg(mt + p);
          *reinterpret_cast<uint4*>(r0 + (p >> 3) * A_BLOCK_BYTES + sw128_offset(r, (p & 7) * 8)) = pk;
        }
#pragma unroll
        for (int p = 0; p < 4; ++p) *reinterpret_cast<uint4*>(r1 + sw128_offset(r, (half * 4 + p) * 8)) = make_uint4(0, 0, 0, 0);
        epi_barrier();  // order the 16-byte fills (both threads of a row) before the element stores below
        if (live && half == 0) {
          const int S = v.spatial;
          for (uint64_t c = v.cheese; c; c &= c - 1) put_elem(r0, A_BLOCK_BYTES, r, 4 * S + (__ffsll((long long)c) - 1), BF16_ONE);
          put_elem(r0, A_BLOCK_BYTES, r, 5 * S, bf16_bits(v.progress));
          put_elem(r1, A_BLOCK_BYTES, r, player ? v.p2 : v.p1, BF16_ONE);
          put_elem(r1, A_BLOCK_BYTES, r, S, bf16_bits(player_elem(v, S, player)));
          put_elem(r1, A_BLOCK_BYTES, r, S + 1, bf16_bits(player_elem(v, S + 1, player)));
        }
      }
      signal_a();
      // stage A results: shared -> r0, p -> r1
      mbar_wait(&sh->mma_done, done_phase); done_phase ^= 1; tc_fence_after();
      hidden_epilogue(t_lane, w.bs, r0, r, half);
      hidden_epilogue(t_lane + 256, w.bp, r1, r, half);
      signal_a();
      // stage B result: t -> r0
      mbar_wait(&sh->mma_done, done_phase); done_phase ^= 1; tc_fence_after();
      hidden_epilogue(t_lane, w.bt1, r0, r, half);
      signal_a();
      // stage C result: h -> r1
      mbar_wait(&sh->mma_done, done_phase); done_phase ^= 1; tc_fence_after();
      hidden_epilogue(t_lane, w.bt2, r1, r, half);
      signal_a();
      // stage D: heads.  z[0:5] = Wa_pol.h_i, z[5] = Wa_val.h_i, z[6:11] = Wb_pol.h_i, z[11] = Wb_val.h_i
      mbar_wait(&sh->mma_done, done_phase); done_phase ^= 1; tc_fence_after();
      if (half != 0) { tc_fence_before(); continue; }  // the heads are 16 columns: the first thread of a row finishes them
      float z[16];
      tmem_ld16(t_lane, z);
      tc_fence_before();
      float o[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        float zb_partner = __shfl_xor_sync(0xffffffffu, z[6 + j], 1);
        o[j] = z[j] + (z[6 + j] + zb_partner) + __ldg(w.bh + j);
      }
      {
        float m = o[0];
#pragma unroll
        for (int j = 1; j < 5; ++j) m = fmaxf(m, o[j]);
        float e[5], s = 0.0f;
#pragma unroll
        for (int j = 0; j < 5; ++j) { e[j] = expf(o[j] - m); s += e[j]; }
#pragma unroll
        for (int j = 0; j < 5; ++j) o[j] = e[j] / s;
        float x = o[5];
        o[5] = x > 20.0f ? x : log1pf(expf(x));
      }
      float q[6];  // the partner's outputs (player 2's, as seen from the even lane)
#pragma unroll
      for (int j = 0; j < 6; ++j) q[j] = __shfl_xor_sync(0xffffffffu, o[j], 1);
      if (live && player == 0) {
        bool ok = true;
#pragma unroll
        for (int j = 0; j < 6; ++j) ok = ok && isfinite(o[j]) && isfinite(q[j]);
        if (!ok) atomicCAS(error_flag, 0, (int)AR_ERR_NONFINITE);
        float4* dst = reinterpret_cast<float4*>(out + (size_t)pidx * 12);
        dst[0] = make_float4(o[0], o[1], o[2], o[3]);
        dst[1] = make_float4(o[4], q[0], q[1], q[2]);
        dst[2] = make_float4(q[3], q[4], o[5], q[5]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == EPI_WARPS + 1) tmem_dealloc(tmem, 512);
}

struct Model : LeafEvaluator {
  uint8_t* d_w = nullptr;
  float* d_b = nullptr;
  Weights w{};
  int n_sms = 148;
  size_t smem_bytes = 0;

  ~Model() override { cudaFree(d_w); cudaFree(d_b); }

  int load(const ar_tensor_desc* t, int n, int width, int height, std::string& err) override {
    const int S = width * height;
    const int shared_dim = 5 * S + 1, player_dim = S + 2;
    const ar_tensor_desc* d0 = find_tensor(t, n, "shared_encoder.0.weight");
    if (!d0 || d0->ndim != 2) { err = "shared_encoder.0.weight missing (not a SymmetricMLP state_dict)"; return AR_ERR_INVALID_ARG; }
    // The kernel computes 256 hidden columns per stage.  A narrower model (hidden_dim H < 256) runs zero-padded: the
    // extra units have zero weights and zero bias, are exactly 0 after their ReLU and add exactly 0 downstream.  The
    // concatenations keep the kernel's strides: cat(shared, p_i) lives at columns [0, H) and [256, 256 + H).
    const int H = (int)d0->shape[0];
    if (H < 1 || H > 256) { err = "hidden_dim must be in [1, 256] for the fused SymmetricMLP kernel (narrower models are zero-padded to its 256 columns), got " + std::to_string(H); return AR_ERR_UNSUPPORTED; }
    if (d0->shape[1] != shared_dim) { err = "shared_encoder.0.weight does not match the board size"; return AR_ERR_INVALID_ARG; }
    const int ks = (shared_dim + KB - 1) / KB;
    if (ks > 4 || player_dim > KB) { err = "board too large for the fused SymmetricMLP kernel (needs 5S+1 <= 256, S+2 <= 64)"; return AR_ERR_UNSUPPORTED; }
    std::vector<float> Wsh, bsh, Wph, bph, Wt1h, bt1h, Wt2h, bt2h, Wpol, bpol, Wval, bval;
    if (!fold_linear_bn(t, n, "shared_encoder.0", "shared_encoder.1", H, shared_dim, Wsh, bsh, err)) return AR_ERR_INVALID_ARG;
    if (!fold_linear_bn(t, n, "player_encoder.0", "player_encoder.1", H, player_dim, Wph, bph, err)) return AR_ERR_INVALID_ARG;
    if (!fold_linear_bn(t, n, "trunk.0", "trunk.1", H, 2 * H, Wt1h, bt1h, err)) return AR_ERR_INVALID_ARG;
    if (!fold_linear_bn(t, n, "trunk.4", "trunk.5", H, H, Wt2h, bt2h, err)) return AR_ERR_INVALID_ARG;
    if (!fold_linear_bn(t, n, "policy_head", "", 5, 2 * H, Wpol, bpol, err)) return AR_ERR_INVALID_ARG;
    if (!fold_linear_bn(t, n, "value_head", "", 1, 2 * H, Wval, bval, err)) return AR_ERR_INVALID_ARG;
    std::vector<float> Ws((size_t)256 * shared_dim, 0.0f), bs(256, 0.0f), Wp((size_t)256 * player_dim, 0.0f), bp(256, 0.0f);
    std::vector<float> Wt1((size_t)256 * 512, 0.0f), bt1(256, 0.0f), Wt2((size_t)256 * 256, 0.0f), bt2(256, 0.0f);
    memcpy(Ws.data(), Wsh.data(), (size_t)H * shared_dim * 4);
    memcpy(Wp.data(), Wph.data(), (size_t)H * player_dim * 4);
    memcpy(bs.data(), bsh.data(), (size_t)H * 4);
    memcpy(bp.data(), bph.data(), (size_t)H * 4);
    memcpy(bt1.data(), bt1h.data(), (size_t)H * 4);
    memcpy(bt2.data(), bt2h.data(), (size_t)H * 4);
    for (int r = 0; r < H; ++r) {
      memcpy(&Wt1[(size_t)r * 512], &Wt1h[(size_t)r * 2 * H], (size_t)H * 4);
      memcpy(&Wt1[(size_t)r * 512 + 256], &Wt1h[(size_t)r * 2 * H + H], (size_t)H * 4);
      memcpy(&Wt2[(size_t)r * 256], &Wt2h[(size_t)r * H], (size_t)H * 4);
    }
    std::vector<float> Wh((size_t)16 * 256, 0.0f);
    for (int j = 0; j < 5; ++j) {
      memcpy(&Wh[(size_t)j * 256], &Wpol[(size_t)j * 2 * H], (size_t)H * 4);
      memcpy(&Wh[(size_t)(6 + j) * 256], &Wpol[(size_t)j * 2 * H + H], (size_t)H * 4);
    }
    memcpy(&Wh[(size_t)5 * 256], &Wval[0], (size_t)H * 4);
    memcpy(&Wh[(size_t)11 * 256], &Wval[H], (size_t)H * 4);
    std::vector<uint8_t> img = swizzled_image(Ws, 256, shared_dim, 256, ks);
    const size_t o_wp = img.size();
    std::vector<uint8_t> i2 = swizzled_image(Wp, 256, player_dim, 256, 1);
    img.insert(img.end(), i2.begin(), i2.end());
    const size_t o_wt1 = img.size();
    i2 = swizzled_image(Wt1, 256, 512, 256, 8);
    img.insert(img.end(), i2.begin(), i2.end());
    const size_t o_wt2 = img.size();
    i2 = swizzled_image(Wt2, 256, 256, 256, 4);
    img.insert(img.end(), i2.begin(), i2.end());
    const size_t o_wh = img.size();
    i2 = swizzled_image(Wh, 16, 256, 16, 4);
    img.insert(img.end(), i2.begin(), i2.end());
    std::vector<float> ball(4 * 256 + 8, 0.0f);
    memcpy(&ball[0], bs.data(), 1024);
    memcpy(&ball[256], bp.data(), 1024);
    memcpy(&ball[512], bt1.data(), 1024);
    memcpy(&ball[768], bt2.data(), 1024);
    for (int j = 0; j < 5; ++j) ball[1024 + j] = bpol[j];
    ball[1024 + 5] = bval[0];
#define CKN(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(_e); return AR_ERR_CUDA; } } while (0)
    CKN(cudaMalloc(&d_w, img.size()));
    CKN(cudaMalloc(&d_b, ball.size() * 4));
    CKN(cudaMemcpy(d_w, img.data(), img.size(), cudaMemcpyHostToDevice));
    CKN(cudaMemcpy(d_b, ball.data(), ball.size() * 4, cudaMemcpyHostToDevice));
    w.ws = d_w; w.wp = d_w + o_wp; w.wt1 = d_w + o_wt1; w.wt2 = d_w + o_wt2; w.wh = d_w + o_wh;
    w.bs = d_b; w.bp = d_b + 256; w.bt1 = d_b + 512; w.bt2 = d_b + 768; w.bh = d_b + 1024;
    w.ks_blocks = ks;
    smem_bytes = (size_t)8 * A_BLOCK_BYTES + N_STAGES * W_STAGE_BYTES + sizeof(Smem) + 1024;
    CKN(cudaFuncSetAttribute(symmetric_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    int dev = 0;
    CKN(cudaGetDevice(&dev));
    CKN(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
#undef CKN
    return AR_OK;
  }

  cudaError_t forward(const EvalRow* rows, const uint32_t* n_rows_dev, int n_rows_max, const ar_game_pod* games,
                      const uint16_t* maze_tab, float* out, int* error_flag, cudaStream_t stream) const override {
    if (n_rows_max <= 0) return cudaSuccess;
    int tiles = (n_rows_max + POS_PER_TILE - 1) / POS_PER_TILE;
    int grid = tiles < n_sms ? tiles : n_sms;
    symmetric_forward_kernel<<<grid, THREADS, smem_bytes, stream>>>(rows, n_rows_dev, n_rows_max, games, maze_tab, w,
                                                                    out, error_flag);
    return cudaGetLastError();
  }
};

}  // namespace sym

LeafEvaluator* make_symmetric_evaluator() { return new sym::Model(); }

}  // namespace ar
