// mcts_half.cuh — two game trees per warp (uniform-prior path): each 16-lane half of a warp owns a tree.
//
// Same algorithm, record layout and arithmetic as mcts_device.cuh (one warp per tree); what changes is the
// mapping of the 256-byte node record onto lanes.  A half-warp lane hl (0..15) holds TWO 8-byte slots of the
// record: slot hl (A: the 10 edges and the priors, exactly the lanes the selection arithmetic already used)
// and slot hl + 16 (B: node value, total visits, links, child table).  Every collective is a width-16 (or
// width-8) shuffle or a ballot, so the two halves never need each other: the SIMT hardware issues an
// instruction once for both trees whenever they are at the same point of the code (the per-level selection,
// the game step, the backup), and serialises them only where they differ.
//
// Collective masks.  Every branch in this file is uniform inside a half, so the lanes executing a collective
// are always one whole half or both.  The member mask of a collective is therefore the mask of the lanes
// that are executing it, __activemask(), taken in the same straight-line region (no half-divergent branch
// between the two): 0xffffffff when the halves run together — one SHFL / VOTE serves both trees — and the
// half's own mask when it runs alone.  (A per-half constant mask is legal too, but ptxas then guards every
// collective with a uniformity test and replays it once per distinct mask through WARPSYNC.COLLECTIVE
// whenever the halves ARE together, the case this kernel exists for: 20 % of all issued instructions in the
// first version, profiles/r2_summary.md §3b.)  Such a mask can never name an absent lane, so a collective
// cannot hang; maxima are width-8 butterflies (REDUX would fold both trees into one value).
//
// Why: the warp-per-tree kernel is instruction-issue and latency bound (1368 warp instructions per
// simulation with 10 of 32 lanes holding an outcome; throughput = resident trees / 8 µs and the register file
// caps the trees at 4736).  Two trees per warp double the resident trees at the same register budget and let
// the hot instructions issue once for two trees.  The kernel's outer control flow is flattened to one
// simulate_batch per loop iteration so that the halves re-converge at every batch instead of drifting apart
// over games of different lengths.
//
// Follows the same reference code as mcts_device.cuh (search.rs:362-1177, tree.rs:52-365, selfplay.rs:415-598).
#pragma once
#include "mcts_device.cuh"

namespace ar {
namespace hw {

constexpr int HB_V = 0, HB_TV = 1, HB_LINKS = 2, HB_CHILD = 3;  // B-slot owner lanes: slot 16 + hl

// Shared memory of one half: [move table 64 x 5 x 2 = 640][maze costs 256][tp: bc x 8][path: max_depth x 4]
// [pend: bc x 32][cstack: (bc + 1) x 8] — the layout of mcts_device.cuh with a five-wide move table (a cell has at
// most five outcomes) and without the leaf states of the NN-guided kernel: 1888 instead of 2520 bytes at batch 16 /
// 50 turns.  With the per-block reserve that takes the one-warp blocks of a streaming launch from 200 KB to 159 KB
// per SM, i.e. from the 228 KB shared-memory carve-out to the 164 KB one: 92 KB of L1 instead of 28 (+1.4 % in an
// A/B; loading the records around the L1 costs 5 %, profiles/r2_half_engine_experiments.log).
constexpr int H_STRIDE = 5, H_MAZE = 640, H_TP = 896;
__host__ __device__ inline size_t half_smem_bytes(uint32_t max_depth, uint32_t batch_cap) {
  size_t b = H_TP + (size_t)batch_cap * 8;
  b += ((size_t)max_depth * 4 + 15) & ~(size_t)15;
  b += (size_t)batch_cap * 32;
  b += (size_t)(batch_cap + 1) * 8;
  return (b + 15) & ~(size_t)15;
}

struct HalfCtx : WarpCtx {
  int hbase;             // 0 or 16: first lane of this half
  const float* fpu_tab;  // shared memory: sqrt of the visited prior mass under uniform priors, [n * 6 + k]
  // these hide the WarpCtx accessors (same names, this half's layout)
  __device__ __forceinline__ const uint16_t* steptbl() const { return reinterpret_cast<const uint16_t*>(sm); }
  __device__ __forceinline__ uint8_t* maze() const { return sm + H_MAZE; }
  __device__ __forceinline__ TpEntry* tp() const { return reinterpret_cast<TpEntry*>(sm + H_TP); }
  __device__ __forceinline__ PendLevel* pend() const {
    return reinterpret_cast<PendLevel*>(reinterpret_cast<uint8_t*>(path) + (((size_t)max_depth * 4 + 15) & ~(size_t)15));
  }
  __device__ __forceinline__ ChildEnt* cstack() const { return reinterpret_cast<ChildEnt*>(pend() + batch_cap); }
  __device__ __forceinline__ void bind_half(uint8_t* base, NodeRec* pool_, int lane, uint32_t max_depth_, uint32_t batch_cap_) {
    bind(base, pool_, lane, max_depth_, batch_cap_);
    path = reinterpret_cast<uint32_t*>(sm + H_TP + (size_t)batch_cap_ * 8);
    AR_OPAQUE(path);
    __builtin_assume(__isShared(path));
  }
#ifdef AR_HALF_IDLE
  unsigned long long idle_pick, idle_backup;  // measurement build: cycles spent waiting for the other half
#endif
};

// sqrt(visited_prior_mass) of compute_fpu (search.rs:120-128) when every prior of the half is 1 / n: the sum
// of the k visited priors in outcome order is k additions of the same f32, whichever outcomes they are.
__device__ __forceinline__ float fpu_tab_entry(int n, int k) {
  float mass = 0.0f;
  if (n > 0) {
    const float p = 1.0f / (float)n;
    for (int i = 0; i < k; ++i) mass = mass + p;
  }
  return sqrtf(mass);
}

// -DAR_HALF_CHECK: trap when a collective is reached by anything but one whole half or both (the invariant the
// __activemask() member masks rest on); the GPU parity suite is run once with this build (profiles/).
__device__ __forceinline__ void check_mask(unsigned am) {
#ifdef AR_HALF_CHECK
  unsigned lane;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
  if (am != FULL && am != (0xffffu << (lane & 16))) {
    printf("half engine: collective reached with lane mask %08x (lane %u)\n", am, lane);
    __trap();
  }
#endif
}
__device__ __forceinline__ uint32_t hshfl(unsigned am, uint32_t v, int src) { check_mask(am); return __shfl_sync(am, v, src, 16); }
__device__ __forceinline__ float hshfl(unsigned am, float v, int src) { check_mask(am); return __shfl_sync(am, v, src, 16); }
__device__ __forceinline__ uint32_t ballot16(unsigned am, const HalfCtx& cx, bool pred) {
  check_mask(am);
  return (__ballot_sync(am, pred) >> cx.hbase) & 0xffffu;
}
__device__ __forceinline__ void hsync() {
  const unsigned am = __activemask();
  check_mask(am);
  __syncwarp(am);
}
// maximum over the lane's 8-lane segment (P1 or P2 outcomes of its tree); scores are never NaN and a zero is
// always +0.0 (the callers add 0.0f), so the float maximum orders them like the reference's comparisons
__device__ __forceinline__ float seg_max(unsigned am, float x) {
  check_mask(am);
  x = fmaxf(x, __shfl_xor_sync(am, x, 1, 8));
  x = fmaxf(x, __shfl_xor_sync(am, x, 2, 8));
  x = fmaxf(x, __shfl_xor_sync(am, x, 4, 8));
  return x;
}
// prior of the lane's outcome: 1 / n under uniform priors (fpu_tab[36 + n], the value write_new_node stored);
// the stored priors only at a root that carries Dirichlet noise
__device__ __forceinline__ float half_prior(const HalfCtx& cx, bool noisy_root, uint2 A, bool valid, int nseg, int seg,
                                            int o) {
  if (!noisy_root) return valid ? cx.fpu_tab[36 + nseg] : 0.0f;
  const unsigned am = __activemask();
  const int psrc = seg + LANE_PRIOR + ((o < 5 ? o : 0) >> 1);
  const uint32_t px = hshfl(am, A.x, psrc), py = hshfl(am, A.y, psrc);
  return valid ? __uint_as_float((o & 1) ? py : px) : 0.0f;
}

__device__ __forceinline__ void load_rec2(const HalfCtx& cx, uint32_t node, uint2& A, uint2& B) {
  A = cx.pool_lane[(size_t)node * 32];
  B = cx.pool_lane[(size_t)node * 32 + 16];
}

// Fresh record (write_new_node of mcts_device.cuh on two slots per lane).
__device__ __forceinline__ void write_new_node(NodeRec* pool, uint32_t idx, uint32_t parent, uint32_t meta,
                                               bool prior_uniform, int hl) {
  uint2 a = make_uint2(0, 0), b = make_uint2(0, 0);
  const int seg = hl & 8, o = hl & 7;
  if (o >= LANE_PRIOR && prior_uniform) {
    const int n = __popc(seg ? meta_m2(meta) : meta_m1(meta));
    const uint32_t p = __float_as_uint(1.0f / (float)n);
    const int o0 = (o - LANE_PRIOR) * 2;
    if (o0 < n) a.x = p;
    if (o0 + 1 < n) a.y = p;
  }
  if (hl == HB_LINKS) {
    b.x = parent;
    b.y = meta;
  }
  pool[idx].s[hl] = a;
  pool[idx].s[hl + 16] = b;
}

// compute_fpu (search.rs:120-128) for the lane's player.  `unv` = the half's ballot of outcomes without
// visits.  Uniform priors (every node but a root that carries Dirichlet noise): table lookup; otherwise the
// sequential sum over the visited outcomes as a lane chain.
__device__ __forceinline__ float half_fpu(const HalfCtx& cx, const SearchParams& sp, bool noisy_root, uint32_t unv,
                                          bool valid, uint32_t visits, float prior, float nodeval, float scale,
                                          int nseg, int seg, int o) {
  if (!noisy_root) {
    const int k = nseg - __popc((unv >> seg) & 0x1fu);
    return nodeval - sp.fpu_reduction * scale * cx.fpu_tab[nseg * 6 + k];
  }
  const unsigned am = __activemask();
  float mass = (valid && visits > 0) ? prior : 0.0f;
#pragma unroll
  for (int i = 1; i < 5; ++i) {
    float up = __shfl_up_sync(am, mass, 1, 16);
    if (o == i) mass = up + mass;
  }
  mass = hshfl(am, mass, seg + 4);
  return nodeval - sp.fpu_reduction * scale * sqrtf(mass);
}

// reservoir sampling over the tied outcomes, P1 then P2 (RNG order, search.rs:779-786); no collectives
__device__ __forceinline__ void break_ties(Rng& rng, uint32_t tie, int& b1, int& b2) {
  uint32_t t1 = tie & 0x1fu, t2 = (tie >> 8) & 0x1fu, tc = 1;
  while (t1) {
    int i = __ffs(t1) - 1;
    t1 &= t1 - 1;
    tc += 1;
    if (rng_gen_range(rng, tc) == 0) b1 = i;
  }
  tc = 1;
  while (t2) {
    int i = __ffs(t2) - 1;
    t2 &= t2 - 1;
    tc += 1;
    if (rng_gen_range(rng, tc) == 0) b2 = i;
  }
}

// build_gather_level for one visit (select_single of mcts_device.cuh)
__device__ __forceinline__ int select_single(HalfCtx& cx, const SearchParams& sp, Rng& rng, uint32_t node,
                                             uint2 A, uint2 B, uint32_t meta, uint32_t tv, bool is_root, int hl,
                                             uint32_t& child_out) {
  const float NEG_INF = __int_as_float(0xff800000);
  const unsigned am = __activemask();
  const int seg = hl & 8, o = hl & 7;
  const int nseg = __popc(seg ? meta_m2(meta) : meta_m1(meta));
  const bool valid = o < nseg;
  const bool noisy_root = is_root && sp.noise_epsilon > 0.0f;
  const uint32_t v1u = hshfl(am, B.x, HB_V), v2u = hshfl(am, B.y, HB_V);
  const uint32_t unv = ballot16(am, cx, valid && (A.y & VIS_MASK) == 0);
  const float prior = half_prior(cx, noisy_root, A, valid, nseg, seg, o);
  const float q = valid ? __uint_as_float(A.x) : 0.0f;
  const uint32_t visits = valid ? (A.y & VIS_MASK) : 0u;
  const uint32_t nif = valid ? (A.y >> VIS_BITS) : 0u;
  const float scale = (float)meta_scale(meta);
  const uint32_t cv = tv > 0 ? tv - 1 : 0;
  const float fpu = half_fpu(cx, sp, noisy_root, unv, valid, visits, prior, __uint_as_float(seg ? v2u : v1u), scale,
                             nseg, seg, o);
  const float sqrt_total = sqrt_count<true>((float)(cv > 1u ? cv : 1u));
  const float qv = visits > 0 ? q : fpu;
  const float q_norm = div_guard<true>(qv, scale);
  const float explo_num = sp.c_puct * prior * sqrt_total;
  const uint32_t ns = visits + nif;
  float score = (q_norm + div_guard<true>(explo_num, 1.0f + (float)ns)) + 0.0f;
  if (is_root && sp.force_k > 0.0f) {
    const bool forced = prior > 0.0f && (float)visits < sqrtf(sp.force_k * prior * (float)cv);
    score = forced ? 1e20f : score;
  }
  score = valid ? score : NEG_INF;
  const unsigned am2 = __activemask();
  const float m = seg_max(am2, score);
  const uint32_t eq = ballot16(am2, cx, valid && score == m);
  int b1 = __ffs(eq & 0x1fu) - 1, b2 = __ffs((eq >> 8) & 0x1fu) - 1;  // first strict max
  const int first = seg ? b2 : b1;
  const uint32_t tie = ballot16(am2, cx, valid && o != first && fabsf(score - m) < 1e-12f);
  if (tie) break_ties(rng, tie, b1, b2);
  const bool mine = valid && o == (seg ? b2 : b1);
  if (mine) cx.pool_lane[(size_t)node * 32].y = visits | ((nif + 1u) << VIS_BITS);
  const int f = b1 * 5 + b2;
  child_out = hshfl(__activemask(), (f & 1) ? B.y : B.x, HB_CHILD + (f >> 1));
  return f;
}

// build_gather_level (search.rs:742-817).  Cells f = hl (vtp_a / child_a) and f = hl + 16 (vtp_b / child_b,
// hl < 9) live on lane hl; returns the 25-bit mask of cells that received visits.
__device__ __forceinline__ uint32_t build_level(HalfCtx& cx, const SearchParams& sp, Rng& rng, uint32_t node,
                                                uint2 A, uint2 B, uint32_t meta, uint32_t cur_limit, bool is_root,
                                                int hl, uint32_t& vtp_a, uint32_t& vtp_b, uint32_t& child_a,
                                                uint32_t& child_b) {
  const float NEG_INF = __int_as_float(0xff800000);
  const unsigned am = __activemask();
  const float v1 = __uint_as_float(hshfl(am, B.x, HB_V)), v2 = __uint_as_float(hshfl(am, B.y, HB_V));
  const uint32_t tv = hshfl(am, B.x, HB_TV);
  const int n1 = __popc(meta_m1(meta)), n2 = __popc(meta_m2(meta));
  const float scale = (float)meta_scale(meta);
  const uint32_t cv = tv > 0 ? tv - 1 : 0;
  const int seg = hl & 8, o = hl & 7;
  const int nseg = seg ? n2 : n1;
  const bool valid = o < nseg;
  const bool noisy_root = is_root && sp.noise_epsilon > 0.0f;
  const uint32_t unv = ballot16(am, cx, valid && (A.y & VIS_MASK) == 0);
  const float prior = half_prior(cx, noisy_root, A, valid, nseg, seg, o);
  const float q = valid ? __uint_as_float(A.x) : 0.0f;
  const uint32_t visits = valid ? (A.y & VIS_MASK) : 0u;
  const uint32_t nif = valid ? (A.y >> VIS_BITS) : 0u;
  const float nodeval = seg ? v2 : v1;
  const float fpu = half_fpu(cx, sp, noisy_root, unv, valid, visits, prior, nodeval, scale, nseg, seg, o);
  const float sqrt_total = sqrt_count<true>((float)(cv > 1u ? cv : 1u));
  const float qv = visits > 0 ? q : fpu;
  const float q_norm = div_guard<true>(qv, scale);
  const float explo_num = sp.c_puct * prior * sqrt_total;
  bool forced = false;
  if (is_root && sp.force_k > 0.0f && prior > 0.0f) {
    const float threshold = sqrtf(sp.force_k * prior * (float)cv);
    forced = (float)visits < threshold;
  }
  uint32_t ns = visits + nif;
  const uint32_t ns0 = ns;
  uint32_t remaining = cur_limit;
  uint32_t va = 0, vb = 0;

  while (remaining > 0) {
    float score = NEG_INF;
    if (valid) score = (forced ? 1e20f : q_norm + div_guard<true>(explo_num, 1.0f + (float)ns)) + 0.0f;
    const unsigned am1 = __activemask();
    const float m = seg_max(am1, score);
    const uint32_t eq = ballot16(am1, cx, valid && score == m);
    const int first1 = __ffs(eq & 0x1fu) - 1, first2 = __ffs((eq >> 8) & 0x1fu) - 1;
    const int first = seg ? first2 : first1;
    const uint32_t tie = ballot16(am1, cx, valid && o != first && fabsf(score - m) < 1e-12f);
    int b1 = first1, b2 = first2;
    if (tie) break_ties(rng, tie, b1, b2);
    const int best = seg ? b2 : b1;
    uint32_t k = 1;
    if (remaining > 1) {
      const unsigned am2 = __activemask();
      const float second = seg_max(am2, (valid && o != first) ? score : NEG_INF);  // NEG_INF: no other outcome
      const float util = hshfl(am2, q_norm, seg + best);
      const float prior_best = hshfl(am2, prior, seg + best);
      const uint32_t ns_best = hshfl(am2, ns, seg + best);
      uint32_t vtc = 0xffffffffu;
      if (!(second <= NEG_INF) && !(util >= second)) {
        const float denom = second - util;
        if (!(denom <= 0.0f)) {
          const float n1f = (float)ns_best + 1.0f;
          const float x = fmaxf(sp.c_puct * prior_best * sqrt_total / denom - n1f + 1.0f, 1.0f);
          const uint32_t u = f2u_sat(x);
          vtc = u > 1u ? u : 1u;
        }
      }
      const uint32_t vto = __shfl_xor_sync(__activemask(), vtc, 8, 16);
      k = vtc < vto ? vtc : vto;
      k = remaining < k ? remaining : k;
      k = k > 1u ? k : 1u;
    }
    if (o == best) ns += k;
    const int f = b1 * 5 + b2;
    if ((f & 15) == hl) {
      if (f < 16) va += k; else vb += k;
    }
    remaining -= k;
  }
  const uint32_t delta = ns - ns0;
  if (valid && delta > 0) cx.pool_lane[(size_t)node * 32].y = visits | ((nif + delta) << VIS_BITS);

  // child[f] sits in slot 19 + f / 2 = B of lane 3 + f / 2, component f & 1
  const unsigned am3 = __activemask();
  const uint32_t ax = hshfl(am3, B.x, HB_CHILD + (hl >> 1)), ay = hshfl(am3, B.y, HB_CHILD + (hl >> 1));
  const int sb = HB_CHILD + 8 + ((hl < 9 ? hl : 0) >> 1);
  const uint32_t bx = hshfl(am3, B.x, sb), by = hshfl(am3, B.y, sb);
  child_a = (hl & 1) ? ay : ax;
  child_b = (hl & 1) ? by : bx;
  vtp_a = va;
  vtp_b = vb;
  return ballot16(am3, cx, va > 0) | (ballot16(am3, cx, hl < 9 && vb > 0) << 16);
}

__device__ __forceinline__ void save_path(HalfCtx& cx, int entry, int depth, uint32_t leaf, int hl) {
  hsync();  // lane 0 wrote cx.path[0..depth) level by level; one barrier here instead of one per level
  uint32_t* pb = cx.path_buf + (size_t)entry * cx.path_stride;
  for (int j = hl; j <= depth; j += 16) pb[j] = j < depth ? cx.path[j] : leaf;
}

// pick_nodes_to_extend (search.rs:576-738), uniform priors (pick_nodes<false> of mcts_device.cuh)
__device__ __forceinline__ uint32_t pick_nodes(HalfCtx& cx, const SearchParams& sp, Rng& rng, const GState& root_g,
                                               int root_turn, uint32_t budget, int& n_tp, int hl) {
  uint32_t collisions = 0;
  uint2 A, B;
  load_rec2(cx, 0, A, B);
  const unsigned am0 = __activemask();
  const uint32_t rtv = hshfl(am0, B.x, HB_TV);
  uint32_t meta = hshfl(am0, B.y, HB_LINKS);
  const bool rterm = meta_term(meta);
  if (rtv == 0 || rterm) {
    const bool over = rterm || game_over(root_g, root_turn, cx.max_turns);
    const bool claim_ok = rtv > 0 || !cx.root_claimed;
    if (claim_ok) {
      cx.root_claimed = true;
      if (over && !rterm && hl == HB_LINKS) cx.pool[0].s[LANE_LINKS].y = meta | (1u << 6);
      if (hl == 0) cx.tp()[n_tp] = TpEntry{0u, (uint8_t)(over ? 1 : 0), 0, 0};
      save_path(cx, n_tp, 0, 0, hl);
      n_tp += 1;
      collisions += budget - 1;
    } else {
      collisions += budget;
    }
    hsync();
    return collisions;
  }

  uint32_t node = 0;
  GState g = root_g;
  int d = 0;
  uint32_t cur_limit = budget, cur_tv = rtv;
  bool is_root = true;
  int n_pend = 0;
  int cs_top = 0;
  for (;;) {
    int m1 = meta_m1(meta), m2 = meta_m2(meta);
    int f;
    uint32_t k, child, rest = 0, va = 0, vb = 0, ca = 0, cb = 0;
    if (cur_limit == 1) {
      f = select_single(cx, sp, rng, node, A, B, meta, cur_tv, is_root, hl, child);
      k = 1;
    } else {
      const uint32_t pending = build_level(cx, sp, rng, node, A, B, meta, cur_limit, is_root, hl, va, vb, ca, cb);
      f = __ffs(pending) - 1;
      const unsigned am = __activemask();
      const uint32_t ka = hshfl(am, va, f & 15), kb = hshfl(am, vb, f & 15);
      const uint32_t cha = hshfl(am, ca, f & 15), chb = hshfl(am, cb, f & 15);
      k = f < 16 ? ka : kb;
      child = f < 16 ? cha : chb;
      rest = pending & (pending - 1);
    }
    if (rest) {  // the level split its visits: park the remaining cells
      if ((rest >> hl) & 1u) {
        const int pos = cs_top + __popc(rest & ((1u << hl) - 1u));
        cx.cstack()[pos] = ChildEnt{ca, (uint8_t)hl, (uint8_t)va, 0};
      }
      if (hl < 9 && ((rest >> (hl + 16)) & 1u)) {
        const int pos = cs_top + __popc(rest & ((1u << (hl + 16)) - 1u));
        cx.cstack()[pos] = ChildEnt{cb, (uint8_t)(hl + 16), (uint8_t)vb, 0};
      }
      if (hl == 0) {
        PendLevel& P = cx.pend()[n_pend];
        P.g = g_pack(g);
        P.node = node;
        P.cs_begin = (uint16_t)cs_top;
        P.cs_cur = (uint16_t)cs_top;
        P.cs_end = (uint16_t)(cs_top + __popc(rest));
        P.m1 = (uint8_t)m1;
        P.m2 = (uint8_t)m2;
        P.depth = (uint8_t)d;
      }
      cs_top += __popc(rest);
      n_pend += 1;
      hsync();
    }
    for (;;) {
      const int a1 = (f * 13) >> 6, a2 = f - a1 * 5;
      uint2 cA = make_uint2(0, 0), cB = make_uint2(0, 0);
      if (child != 0) load_rec2(cx, child, cA, cB);
      GState gc = g;
      game_step<H_STRIDE>(gc, a1, a2, cx.steptbl());
      const int rc = (gc.s1x2 - g.s1x2) | ((gc.s2x2 - g.s2x2) << 2);
      const int child_turn = root_turn + d + 1;
      if (hl == 0) cx.path[d] = node | ((uint32_t)f << PATH_NODE_BITS) | ((uint32_t)rc << 28);
      bool descend = false;
      if (child == 0) {
        if (cx.node_count >= cx.pool_nodes || n_tp >= MAX_BATCH) {
          cx.error = AR_ERR_POOL_OVERFLOW;
          return collisions;
        }
        child = cx.node_count++;
        cx.new_nodes++;
        const bool over = game_over(gc, child_turn, cx.max_turns);
        const int cm1 = eff_mask(cx.maze(), gc.p1, gc.mud1), cm2 = eff_mask(cx.maze(), gc.p2, gc.mud2);
        const int rem = __popcll(gc.cheese);
        const uint32_t cmeta = meta_pack(a1, a2, over ? 1 : 0, cm1, cm2, rem > 1 ? rem : 1, rc & 3, rc >> 2);
        write_new_node(cx.pool, child, node, cmeta, !over, hl);
        if (hl == 0) {
          reinterpret_cast<uint32_t*>(&cx.pool[node].s[LANE_CHILD])[f] = child;
          cx.tp()[n_tp] = TpEntry{child, (uint8_t)(over ? 1 : 0), (uint8_t)(d + 1), 0};
        }
        save_path(cx, n_tp, d + 1, child, hl);
        n_tp += 1;
        collisions += k - 1;
      } else {
        const unsigned am = __activemask();
        const uint32_t ctv = hshfl(am, cB.x, HB_TV);
        const uint32_t cmeta = hshfl(am, cB.y, HB_LINKS);
        if (ctv == 0) {
          collisions += k;
        } else if (meta_term(cmeta)) {
          if (n_tp >= MAX_BATCH) { cx.error = AR_ERR_POOL_OVERFLOW; return collisions; }
          if (hl == 0) cx.tp()[n_tp] = TpEntry{child, 1, (uint8_t)(d + 1), 0};
          save_path(cx, n_tp, d + 1, child, hl);
          n_tp += 1;
          collisions += k - 1;
        } else {
          if ((uint32_t)(d + 1) >= cx.max_depth) {
            cx.error = AR_ERR_POOL_OVERFLOW;
            return collisions;
          }
          node = child; A = cA; B = cB; meta = cmeta; g = gc; d += 1; cur_limit = k; cur_tv = ctv; is_root = false;
          descend = true;
        }
      }
      if (descend) break;
      if (n_pend == 0) return collisions;
      hsync();
      PendLevel& P = cx.pend()[n_pend - 1];
      const int cur = P.cs_cur, end = P.cs_end;
      const ChildEnt ce = cx.cstack()[cur];
      g = g_unpack(P.g);
      node = P.node; m1 = P.m1; m2 = P.m2; d = P.depth;
      f = ce.f; k = ce.k; child = ce.child;
      hsync();
      if (cur + 1 == end) {
        n_pend -= 1;
        cs_top = P.cs_begin;
      } else if (hl == 0) {
        P.cs_cur = (uint16_t)(cur + 1);  // read again only after the barrier at the top of the next pop
      }
    }
  }
}

// backup_and_finalize (search.rs:826-852), path-parallel over 16 lanes, leaf value 0 (uniform priors): the chain
// q_j = r_j + q_{j+1} is an integer suffix scan in half units (backup_entry<true> of mcts_device.cuh).
__device__ __forceinline__ void backup_entry(HalfCtx& cx, int entry, int hl) {
  const TpEntry te = cx.tp()[entry];
  const int depth = te.depth;
  const uint32_t* pb = cx.path_buf + (size_t)entry * cx.path_stride;
  cx.path_nodes += depth + 1;
  uint32_t carry = 0;
  for (int hi = depth; hi >= 0; hi -= 16) {
    const int lo = hi - 15 > 0 ? hi - 15 : 0;
    const int j = lo + hl;
    const bool active = j <= hi;
    const uint32_t e = active ? pb[j] : 0u;
    const bool is_leaf = active && j == depth;
    const uint32_t node = e & PATH_NODE_MASK;
    const int f = (e >> PATH_NODE_BITS) & 31;
    const int a1 = (f * 13) >> 6, a2 = f - a1 * 5;
    uint4 st = make_uint4(0, 0, 0, 0);
    uint2 e1 = make_uint2(0, 0), e2 = e1;
    if (active) {
      st = *reinterpret_cast<const uint4*>(&cx.pool[node].s[LANE_V]);
      if (!is_leaf) {
        e1 = cx.pool[node].s[a1];
        e2 = cx.pool[node].s[LANE_P2 + a2];
      }
    }
    uint32_t sfx = (active && !is_leaf) ? (((e >> 28) & 3u) | ((e >> 30) << 16)) : 0u;
    const unsigned am = __activemask();
#pragma unroll
    check_mask(am);
    for (int sh = 1; sh < 16; sh <<= 1) {
      const uint32_t y = __shfl_down_sync(am, sfx, sh, 16);
      if (hl + sh < 16) sfx += y;
    }
    sfx += carry;
    carry = hshfl(am, sfx, 0);
    const float q1 = 0.5f * (float)(sfx & 0xffffu);
    const float q2 = 0.5f * (float)(sfx >> 16);
    if (active) {
      const uint32_t tv = st.z + 1;
      const float n = (float)tv;
      float v1 = __uint_as_float(st.x), v2 = __uint_as_float(st.y);
      v1 = v1 + div_guard<true>((q1 - v1) * 1.0f, n);
      v2 = v2 + div_guard<true>((q2 - v2) * 1.0f, n);
      st.x = __float_as_uint(v1);
      st.y = __float_as_uint(v2);
      st.z = tv;
      *reinterpret_cast<uint4*>(&cx.pool[node].s[LANE_V]) = st;
      if (!is_leaf) {
        uint32_t vis = (e1.y & VIS_MASK) + 1;
        float q = __uint_as_float(e1.x);
        q = q + div_guard<true>((q1 - q) * 1.0f, (float)vis);
        e1.x = __float_as_uint(q);
        e1.y = vis;
        cx.pool[node].s[a1] = e1;
        vis = (e2.y & VIS_MASK) + 1;
        q = __uint_as_float(e2.x);
        q = q + div_guard<true>((q2 - q) * 1.0f, (float)vis);
        e2.x = __float_as_uint(q);
        e2.y = vis;
        cx.pool[node].s[LANE_P2 + a2] = e2;
      }
    }
    hsync();
  }
}

// apply_dirichlet_noise (search.rs:400-429); same draws as apply_root_noise of mcts_device.cuh
__device__ __noinline__ void apply_root_noise(HalfCtx& cx, const SearchParams& sp, Rng& rng, int hl) {
  const uint32_t meta = cx.pool[0].s[LANE_LINKS].y;
#pragma unroll 1
  for (int pl = 0; pl < 2; ++pl) {
    const int n = __popc(pl ? meta_m2(meta) : meta_m1(meta));
    if (n <= 1) continue;
    const double alpha = (double)(sp.noise_concentration / (float)n);
    if (!(alpha > 0.0)) continue;
    float noise[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    float total = 0.0f;
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
      const float gm = (float)rng_gamma(rng, alpha);
#pragma unroll
      for (int j = 0; j < 5; ++j)
        if (j == i) noise[j] = gm;
      total = total + gm;
    }
    if (total < 1.17549435e-38f) continue;
    const int base = pl * 8 + LANE_PRIOR;
    if (hl >= base && hl < base + 3) {
      uint2 pr = cx.pool[0].s[hl];
      const int o0 = (hl - base) * 2;
      float p0 = __uint_as_float(pr.x), p1 = __uint_as_float(pr.y);
      float n0 = 0.f, n1 = 0.f;
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        if (j == o0) n0 = noise[j];
        if (j == o0 + 1) n1 = noise[j];
      }
      if (o0 < n) p0 = p0 * (1.0f - sp.noise_epsilon) + sp.noise_epsilon * n0 / total;
      if (o0 + 1 < n) p1 = p1 * (1.0f - sp.noise_epsilon) + sp.noise_epsilon * n1 / total;
      cx.pool[0].s[hl] = make_uint2(__float_as_uint(p0), __float_as_uint(p1));
    }
    hsync();
  }
}

// extract_result (search.rs:1079-1177)
__device__ __forceinline__ void extract_result(HalfCtx& cx, const SearchParams& sp, int hl, ar_search_result& out) {
  uint2 A, B;
  load_rec2(cx, 0, A, B);
  const unsigned am = __activemask();
  const float v1 = __uint_as_float(hshfl(am, B.x, HB_V)), v2 = __uint_as_float(hshfl(am, B.y, HB_V));
  const uint32_t tv = hshfl(am, B.x, HB_TV);
  const uint32_t meta = hshfl(am, B.y, HB_LINKS);
  const float scale = (float)meta_scale(meta);
  const uint32_t cv = tv > 0 ? tv - 1 : 0;
  float pr[2][5], qe[2][5];
  uint32_t vi[2][5];
#pragma unroll
  for (int p = 0; p < 2; ++p)
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const uint32_t px = hshfl(am, A.x, p * 8 + LANE_PRIOR + (i >> 1));
      const uint32_t py = hshfl(am, A.y, p * 8 + LANE_PRIOR + (i >> 1));
      pr[p][i] = __uint_as_float((i & 1) ? py : px);
      qe[p][i] = __uint_as_float(hshfl(am, A.x, p * 8 + i));
      vi[p][i] = hshfl(am, A.y, p * 8 + i) & VIS_MASK;
    }
  extract_half(pr[0], qe[0], vi[0], meta_m1(meta), v1, scale, cv, sp, out.policy_p1, out.visit_counts_p1,
               out.value_p1, out.prior_p1, out.raw_visits_p1);
  extract_half(pr[1], qe[1], vi[1], meta_m2(meta), v2, scale, cv, sp, out.policy_p2, out.visit_counts_p2,
               out.value_p2, out.prior_p2, out.raw_visits_p2);
  out.total_visits = tv;
  out.node_count = cx.node_count;
  out.reserved = 0;
}

__device__ __forceinline__ void init_root(HalfCtx& cx, const GState& g, int hl) {
  const int m1 = eff_mask(cx.maze(), g.p1, g.mud1), m2 = eff_mask(cx.maze(), g.p2, g.mud2);
  const int rem = __popcll(g.cheese);
  const uint32_t meta = meta_pack(0, 0, 0, m1, m2, rem > 1 ? rem : 1, 0, 0);
  write_new_node(cx.pool, 0, NO_PARENT, meta, true, hl);
  cx.node_count = 1;
  hsync();
}

// advance_root with in-place compaction (compact_subtree of mcts_device.cuh), 16 nodes per marking step and
// two records per sliding step
__device__ __forceinline__ void compact_subtree(HalfCtx& cx, uint32_t new_root, int hl) {
  const uint32_t count = cx.node_count;
  uint32_t* remap = cx.remap;
  uint32_t kept = 0;
  for (uint32_t base = new_root; base < count; base += 16) {
    const uint32_t node = base + hl;
    const bool in = node < count;
    const uint32_t parent = in ? cx.pool[node].s[LANE_LINKS].x : NO_PARENT;
    bool keep = in && node == new_root;
    const bool local = in && node != new_root && parent != NO_PARENT && parent >= base;
    if (in && node != new_root && parent != NO_PARENT && parent >= new_root && parent < base)
      keep = remap[parent] != NO_PARENT;
    uint32_t km = ballot16(__activemask(), cx, keep);
    for (;;) {
      const bool k2 = keep || (local && ((km >> (parent - base)) & 1u));
      const uint32_t nm = ballot16(__activemask(), cx, k2);
      keep = k2;
      if (nm == km) break;
      km = nm;
    }
    const uint32_t rank = kept + __popc(km & ((1u << hl) - 1u));
    if (in) remap[node] = keep ? rank : NO_PARENT;
    kept += __popc(km);
    hsync();
  }
  uint32_t next = new_root;
  while (next < count) {
    const uint32_t cand = next + hl;
    const uint32_t cdst = cand < count ? remap[cand] : NO_PARENT;
    uint32_t cm = ballot16(__activemask(), cx, cdst != NO_PARENT);
    if (cm == 0) { next += 16; continue; }
    uint32_t src[2], dst[2];
    uint2 va[2], vb[2];
    int last = 0;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      if (cm) {
        const int i = __ffs(cm) - 1;
        cm &= cm - 1;
        src[t] = next + i;
        dst[t] = hshfl(__activemask(), cdst, i);
        last = i;
      } else {
        src[t] = NO_PARENT;
        dst[t] = hshfl(__activemask(), cdst, 0);
      }
    }
#pragma unroll
    for (int t = 0; t < 2; ++t)
      if (src[t] != NO_PARENT) {
        va[t] = cx.pool[src[t]].s[hl];
        vb[t] = cx.pool[src[t]].s[hl + 16];
      }
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      if (src[t] == NO_PARENT) continue;
      if (hl == HB_LINKS) vb[t].x = (src[t] == new_root) ? NO_PARENT : remap[vb[t].x];
      if (hl >= HB_CHILD) {
        if (vb[t].x) vb[t].x = remap[vb[t].x];
        if (vb[t].y) vb[t].y = remap[vb[t].y];
      }
    }
    hsync();
#pragma unroll
    for (int t = 0; t < 2; ++t)
      if (src[t] != NO_PARENT) {
        cx.pool[dst[t]].s[hl] = va[t];
        cx.pool[dst[t]].s[hl + 16] = vb[t];
      }
    hsync();
    next = next + last + 1;
  }
  cx.node_count = kept;
}

// One simulate_batch (search.rs:961-1073) with SmartUniformBackend fused in.
__device__ __forceinline__ void simulate_batch_uniform(HalfCtx& cx, const SearchParams& sp, Rng& rng,
                                                       const GState& root_g, int root_turn, uint32_t bs, uint32_t& nn,
                                                       uint32_t& term, uint32_t& coll, uint32_t coll_len, int hl) {
  cx.root_claimed = false;
  const uint32_t ci = cx.node_count < coll_len ? cx.node_count : coll_len - 1;
  int collisions_left = (int)cx.coll_table[ci];
  int n_tp = 0;
#ifdef AR_HALF_IDLE
  long long t_last = clock64();
#endif
  while ((uint32_t)n_tp < bs && collisions_left > 0 && cx.error == 0) {
    const uint32_t budget = min((uint32_t)collisions_left, bs - (uint32_t)n_tp);
    const uint32_t c = pick_nodes(cx, sp, rng, root_g, root_turn, budget, n_tp, hl);
    collisions_left -= (int)c;
    coll += c;
#ifdef AR_HALF_IDLE
    t_last = clock64();
#endif
  }
#ifdef AR_HALF_IDLE
  __syncwarp(__activemask());
  cx.idle_pick += clock64() - t_last;
  t_last = clock64();
#endif
  if (cx.error) return;
  for (int e = 0; e < n_tp; ++e) {
    const uint8_t kind = cx.tp()[e].kind;
    if (kind == 1) term += 1; else nn += 1;
    if (kind == 0 && cx.tp()[e].node == 0 && sp.noise_epsilon > 0.0f) apply_root_noise(cx, sp, rng, hl);
    backup_entry(cx, e, hl);
#ifdef AR_HALF_IDLE
    t_last = clock64();
#endif
  }
#ifdef AR_HALF_IDLE
  __syncwarp(__activemask());
  cx.idle_backup += clock64() - t_last;
#endif
}

__device__ __forceinline__ void load_game(const ar_game_pod* pod, HalfCtx& cx, GState& g, int& turn, int hl) {
  cx.w = pod->width;
  cx.cells = (int)pod->width * pod->height;
  cx.max_turns = pod->max_turns;
  turn = pod->turn;
  hsync();
  for (int i = hl; i < 64; i += 16)
    reinterpret_cast<uint32_t*>(cx.maze())[i] =
        (i < cx.cells) ? reinterpret_cast<const uint32_t*>(pod->move_cost)[i] : 0u;
  uint16_t* tbl = const_cast<uint16_t*>(cx.steptbl());
  for (int i = hl; i < 64 * H_STRIDE; i += 16) {
    const int c = i / H_STRIDE, oi = i - c * H_STRIDE;
    uint32_t e = (uint32_t)c;
    if (c < cx.cells) {
      int a = -1, seen = 0;
#pragma unroll
      for (int dd = 0; dd < 4; ++dd)
        if (pod->move_cost[c * 4 + dd] != 0) {
          if (seen == oi) a = dd;
          seen += 1;
        }
      if (a >= 0) {
        const int cost = pod->move_cost[c * 4 + a];
        const int mag = (a & 1) ? 1 : cx.w;
        e = (uint32_t)(c + ((a & 2) ? -mag : mag)) | ((uint32_t)(cost >= 2 ? cost : 0) << 8);
      }
    }
    tbl[i] = (uint16_t)e;
  }
  g.cheese = *reinterpret_cast<const uint64_t*>(pod->cheese);
  g.p1 = pod->p1_y * pod->width + pod->p1_x;
  g.p2 = pod->p2_y * pod->width + pod->p2_x;
  g.mud1 = pod->p1_mud;
  g.mud2 = pod->p2_mud;
  g.s1x2 = __float2int_rn(pod->p1_score * 2.0f);
  g.s2x2 = __float2int_rn(pod->p2_score * 2.0f);
  hsync();
}

}  // namespace hw
}  // namespace ar
