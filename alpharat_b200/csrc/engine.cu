// engine.cu — kernels and C-ABI of libalpharat_cuda.so (see include/alpharat_cuda.h).
//
// Compiled with -fmad=false: the tree kernels replay the reference's f32 arithmetic
// operation-for-operation (search.rs, node.rs) and must not contract a*b+c into FMA.
// The leaf-evaluator kernels live in nn_kernels.cu (separate translation unit, FMA allowed).
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "host_tables.hpp"
#include "mcts_device.cuh"
#include "mcts_half.cuh"
#include "nn_api.cuh"
#include "tree_thread.cuh"

namespace ar {

// ---------------------------------------------------------------------------------------
// Kernel parameters
// ---------------------------------------------------------------------------------------
struct RunParams {
  // inputs (device)
  const ar_game_pod* games;
  const uint64_t* seeds;
  int n_games;
  // outputs (device)
  ar_game_summary* summaries;
  ar_position_record* positions;
  int pos_stride;
  ar_search_result* search_out;  // search-only mode
  int search_only;
  // search
  SearchParams sp;
  // per-slot storage
  NodeRec* pools;
  uint32_t pool_nodes;
  uint32_t* path_bufs;
  uint32_t path_stride;
  uint32_t* remaps;
  const uint16_t* coll_table;
  uint32_t coll_table_len;
  uint32_t max_depth;
  uint32_t batch_cap;
  int n_slots;
  uint32_t* slot_bitmap;  // streaming launches: blocks claim a group of 4 tree slots here (null: group = blockIdx)
  int n_groups;
  // bookkeeping
  unsigned int* next_game;
  unsigned long long* counters;  // [0] path_nodes [1] new_nodes
  int* error_flag;
  ar_progress* progress;  // mapped pinned host memory (may be null)
};

// The uniform-prior self-play kernel: each warp claims games from an atomic counter
// (game_worker_loop, selfplay.rs:609-650) and plays them to completion on device
// (play_game, selfplay.rs:515-598).  search_only: one fresh-tree search per "game"
// (rust_mcts_search, mcts/bindings.rs:228-304).
#ifndef AR_MIN_BLOCKS
#define AR_MIN_BLOCKS 8
#endif
// WPB = warps per block: 4 for a blocking launch; 1 for streaming launches, where a finished warp must give its
// SM resources back at once (in a 4-warp block the three early finishers idle until the last warp's games end).
template <int WPB>
__global__ void __launch_bounds__(32 * WPB, AR_MIN_BLOCKS * 4 / WPB) selfplay_uniform_kernel(RunParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  int lane = threadIdx.x & 31;
  asm volatile("" : "+r"(lane));
  const int wib = threadIdx.x >> 5;
  // Streaming launches overlap in time (batch b+1 fills the SMs as batch b's last games drain), so a
  // block cannot own the slots of its blockIdx: it claims a free group of WPB tree slots instead.
  __shared__ int s_group;
  int group = blockIdx.x;
  if (p.slot_bitmap) {
    if (threadIdx.x == 0) {
      const int words = (p.n_groups + 31) / 32;
      int w = (int)(blockIdx.x % words), got = -1;
      while (got < 0) {
        for (int t = 0; t < words && got < 0; ++t) {
          uint32_t cur = *reinterpret_cast<volatile uint32_t*>(&p.slot_bitmap[w]);
          while (cur != 0xffffffffu) {
            const int b = __ffs((int)~cur) - 1;
            if (w * 32 + b >= p.n_groups) break;
            const uint32_t old = atomicOr(&p.slot_bitmap[w], 1u << b);
            if (!(old & (1u << b))) { got = w * 32 + b; break; }
            cur = old | (1u << b);
          }
          w = (w + 1 == words) ? 0 : w + 1;
        }
        if (got < 0) __nanosleep(2000);  // more resident blocks than groups: wait for a block to leave
      }
      s_group = got;
    }
    __syncthreads();
    group = s_group;
  }
  const int slot = group * WPB + wib;
  if (slot >= p.n_slots) return;

  uint8_t* base = smem + (size_t)wib * warp_smem_bytes(p.max_depth, p.batch_cap);
  WarpCtx cx;
  cx.bind(base, p.pools + (size_t)slot * p.pool_nodes, lane, p.max_depth, p.batch_cap);
  cx.path_buf = p.path_bufs + (size_t)slot * p.batch_cap * p.path_stride;
  cx.remap = p.remaps + (size_t)slot * p.pool_nodes;
  cx.coll_table = p.coll_table;
  cx.pool_nodes = p.pool_nodes;
  cx.path_stride = p.path_stride;
  cx.epoch = 1;
  cx.path_nodes = 0;
  cx.new_nodes = 0;
  cx.error = 0;
  cx.node_count = 0;
  cx.root_claimed = false;
#ifdef AR_PHASE_TIMING
  for (int i = 0; i < 4; ++i) cx.phase[i] = 0;
  long long t_begin = clock64();
#endif
  const SearchParams sp = p.sp;

  for (;;) {
    unsigned int gi = 0;
    if (lane == 0) gi = atomicAdd(p.next_game, 1u);
    gi = __shfl_sync(FULL, gi, 0);
    if (gi >= (unsigned)p.n_games) break;

    GState g;
    int turn;
    load_game(p.games + gi, cx, g, turn, lane);
    Rng rng = rng_seed(p.seeds[gi]);
    const int cheese_available = __popcll(g.cheese);
    cx.epoch += 1;
    init_root(cx, g, lane);
    if (!p.search_only) init_cheese_outcomes(p.summaries[gi], lane);

    uint32_t n_pos = 0;
    unsigned long long tot_sims = 0, tot_nn = 0, tot_term = 0, tot_coll = 0;
    ar_position_record* pos = p.positions ? p.positions + (size_t)gi * p.pos_stride : nullptr;

    while (p.search_only || !game_over(g, turn, cx.max_turns)) {
      // ---- run_search (search.rs:362-390)
      uint32_t remaining = sp.n_sims, nn = 0, term = 0, coll = 0;
      while (remaining > 0 && cx.error == 0) {
        uint32_t bs = min(remaining, sp.batch_size);
        uint32_t nn0 = nn, term0 = term;
        simulate_batch_uniform(cx, sp, rng, g, turn, bs, nn, term, coll, p.coll_table_len, lane);
        uint32_t produced = (nn - nn0) + (term - term0);
        produced = produced > 1u ? produced : 1u;
        remaining = remaining > produced ? remaining - produced : 0u;
      }
      if (cx.error) break;

      float pol1[5], pol2[5];
      ar_search_result* rout = p.search_only ? (p.search_out + gi) : &pos[n_pos].search;
      extract_and_store(cx, sp, lane, rout, nn, term, coll, pol1, pol2);
      if (p.search_only) break;

      // ---- one self-play move (selfplay.rs:547-565)
      uint32_t tv = 0;
      if (lane == 0) tv = rout->total_visits;
      tv = __shfl_sync(FULL, tv, 0);
      tot_sims += tv; tot_nn += nn; tot_term += term; tot_coll += coll;
      int a1 = rng_sample_action(rng, pol1);
      int a2 = rng_sample_action(rng, pol2);
      if (lane == 0) {
        ar_position_record& pr = pos[n_pos];
        pr.p1_x = (uint8_t)(g.p1 % cx.w); pr.p1_y = (uint8_t)(g.p1 / cx.w);
        pr.p2_x = (uint8_t)(g.p2 % cx.w); pr.p2_y = (uint8_t)(g.p2 / cx.w);
        pr.p1_mud = (uint8_t)g.mud1; pr.p2_mud = (uint8_t)g.mud2;
        pr.action_p1 = (uint8_t)a1; pr.action_p2 = (uint8_t)a2;
        pr.turn = (uint16_t)turn; pr.reserved = 0;
        pr.p1_score = 0.5f * (float)g.s1x2; pr.p2_score = 0.5f * (float)g.s2x2;
        uint32_t* cb = reinterpret_cast<uint32_t*>(pr.cheese);  // 4-byte aligned only
        cb[0] = (uint32_t)g.cheese; cb[1] = (uint32_t)(g.cheese >> 32);
#pragma unroll
        for (int t = 2; t < 8; ++t) cb[t] = 0;
      }
      n_pos += 1;

      // advance_root maps raw actions through action_to_outcome_idx (tree.rs:283-295)
      uint32_t rmeta = cx.pool[0].s[LANE_LINKS].y;
      int i = action_to_idx(meta_m1(rmeta), a1), j = action_to_idx(meta_m2(rmeta), a2);
      uint32_t child = reinterpret_cast<const uint32_t*>(&cx.pool[0].s[LANE_CHILD])[i * 5 + j];
      const uint64_t cheese_before = g.cheese;
      game_step(g, i, j, cx.steptbl());  // the table is indexed by outcome (blocked move = STAY's outcome)
      turn += 1;
      if (lane == 0) credit_cheese(p.summaries[gi], cheese_before, g);
      __syncwarp();
      AR_T0();
      if (game_over(g, turn, cx.max_turns)) break;  // the tree of a finished game is dropped
      if (child != 0) {
        compact_subtree(cx, child, lane);
        AR_T1(2);
      } else {
        cx.epoch += 1;
        init_root(cx, g, lane);  // reinit, tree.rs:298-302
      }
    }

    if (cx.error) break;
    if (!p.search_only && lane == 0) {
      ar_game_summary& s = p.summaries[gi];
      s.game_index = gi;
      s.n_positions = n_pos;
      s.final_p1_score = 0.5f * (float)g.s1x2;
      s.final_p2_score = 0.5f * (float)g.s2x2;
      s.result = g.s1x2 > g.s2x2 ? 1 : (g.s2x2 > g.s1x2 ? 2 : 0);
      s.cheese_available = (uint16_t)cheese_available;
      s.total_simulations = tot_sims;
      s.total_nn_evals = tot_nn;
      s.total_terminals = tot_term;
      s.total_collisions = tot_coll;
      s.reserved[0] = s.reserved[1] = s.reserved[2] = 0;
      s.reserved1 = 0;
      if (p.progress) {
        atomicAdd_system((unsigned long long*)&p.progress->positions_completed, (unsigned long long)n_pos);
        atomicAdd_system((unsigned long long*)&p.progress->simulations_completed, tot_sims);
        atomicAdd_system((unsigned long long*)&p.progress->nn_evals_completed, tot_nn);
        atomicAdd_system((unsigned int*)&p.progress->games_completed, 1u);
      }
#ifndef AR_PHASE_TIMING
      // run totals, so that a caller that only wants throughput needs no record download
      atomicAdd(&p.counters[2], tot_nn);
      atomicAdd(&p.counters[3], tot_term);
      atomicAdd(&p.counters[4], (unsigned long long)n_pos);
      atomicAdd(&p.counters[5], tot_sims);
#endif
    }
  }
#ifdef AR_PHASE_TIMING
  cx.phase[3] = (unsigned long long)(clock64() - t_begin);
#endif
  if (lane == 0) {
    atomicAdd(&p.counters[0], (unsigned long long)cx.path_nodes);
    atomicAdd(&p.counters[1], (unsigned long long)cx.new_nodes);
#ifdef AR_PHASE_TIMING
    for (int i = 0; i < 4; ++i) atomicAdd(&p.counters[2 + i], cx.phase[i]);
#endif
    if (cx.error) atomicCAS(p.error_flag, 0, (int)cx.error);
  }
  if (p.slot_bitmap) {  // n_slots is a multiple of WPB in streaming mode: every warp of the block gets here
    __syncthreads();
    if (threadIdx.x == 0) atomicAnd(&p.slot_bitmap[group >> 5], ~(1u << (group & 31)));
  }
}


// =========================================================================================
// Uniform-prior self-play / search, TWO trees per warp (mcts_half.cuh): each 16-lane half of a warp plays its
// own games.  The outer control flow is flat — one simulate_batch per loop iteration for whichever state the
// half is in — so the two halves of a warp meet again at every batch and the hot code (per-level selection,
// game step, backup) issues once for both trees.
// =========================================================================================
__device__ __noinline__ void extract_and_store_h(hw::HalfCtx& cx, const SearchParams& sp, int hl, ar_search_result* out,
                                                 uint32_t nn, uint32_t term, uint32_t coll, float pol1[5], float pol2[5],
                                                 uint32_t& total_visits) {
  ar_search_result res;
  hw::extract_result(cx, sp, hl, res);
  res.nn_evals = nn;
  res.terminals = term;
  res.collisions = coll;
#pragma unroll
  for (int a = 0; a < 5; ++a) { pol1[a] = res.policy_p1[a]; pol2[a] = res.policy_p2[a]; }
  total_visits = res.total_visits;
  if (hl == 0) *out = res;
}

template <int WPB>
__global__ void __launch_bounds__(32 * WPB, AR_MIN_BLOCKS * 4 / WPB) selfplay_half_kernel(RunParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  int lane = threadIdx.x & 31;
  asm volatile("" : "+r"(lane));
  int hl = lane & 15;
  const int half = lane >> 4;
  const int wib = threadIdx.x >> 5;
  __shared__ int s_group;
  __shared__ float s_fpu_tab[42];  // [n * 6 + k]: sqrt of the visited prior mass; [36 + n]: the uniform prior 1 / n
  if (threadIdx.x < 32) {
    s_fpu_tab[threadIdx.x] = hw::fpu_tab_entry(threadIdx.x / 6, threadIdx.x % 6);
    if (threadIdx.x < 4) s_fpu_tab[32 + threadIdx.x] = hw::fpu_tab_entry(5, 2 + threadIdx.x);
    if (threadIdx.x < 6) s_fpu_tab[36 + threadIdx.x] = threadIdx.x ? 1.0f / (float)threadIdx.x : 0.0f;
  }
  __syncthreads();
  int group = blockIdx.x;
  if (p.slot_bitmap) {  // streaming launches: claim a free group of 2 * WPB tree slots (see selfplay_uniform_kernel)
    if (threadIdx.x == 0) {
      const int words = (p.n_groups + 31) / 32;
      int w = (int)(blockIdx.x % words), got = -1;
      while (got < 0) {
        for (int t = 0; t < words && got < 0; ++t) {
          uint32_t cur = *reinterpret_cast<volatile uint32_t*>(&p.slot_bitmap[w]);
          while (cur != 0xffffffffu) {
            const int b = __ffs((int)~cur) - 1;
            if (w * 32 + b >= p.n_groups) break;
            const uint32_t old = atomicOr(&p.slot_bitmap[w], 1u << b);
            if (!(old & (1u << b))) { got = w * 32 + b; break; }
            cur = old | (1u << b);
          }
          w = (w + 1 == words) ? 0 : w + 1;
        }
        if (got < 0) __nanosleep(2000);
      }
      s_group = got;
    }
    __syncthreads();
    group = s_group;
  }
  const int slot = (group * WPB + wib) * 2 + half;
  const bool live = slot < p.n_slots;
  const int bslot = live ? slot : 0;  // an idle half binds valid pointers and never uses them

  uint8_t* base = smem + (size_t)(wib * 2 + half) * hw::half_smem_bytes(p.max_depth, p.batch_cap);
  hw::HalfCtx cx;
  {
    const float* ft = s_fpu_tab;
    asm volatile("" : "+l"(ft));
    __builtin_assume(__isShared(ft));
    cx.fpu_tab = ft;
  }
  // The pool of the second half starts 128 bytes into its slot (the allocation carries the slack): both 128-byte
  // parts of a record stay line-aligned, and bits 3..7 of the lane's record pointer — a register that is live
  // throughout — spell the lane.  hl and hbase are taken from there, so wherever the compiler re-derives them
  // under register pressure it is a shift and a mask instead of an S2R round trip (5 % of the stall samples).
  cx.bind_half(base, reinterpret_cast<NodeRec*>(reinterpret_cast<uint8_t*>(p.pools + (size_t)bslot * p.pool_nodes) + 128 * half),
               hl, p.max_depth, p.batch_cap);
  {
    const uint32_t lane_bits = ((uint32_t)(uintptr_t)cx.pool_lane >> 3) & 31u;
    hl = (int)(lane_bits & 15u);
    cx.hbase = (int)(lane_bits & 16u);
  }
  cx.path_buf = p.path_bufs + (size_t)bslot * p.batch_cap * p.path_stride;
  cx.remap = p.remaps + (size_t)bslot * p.pool_nodes;
  cx.coll_table = p.coll_table;
  cx.pool_nodes = p.pool_nodes;
  cx.path_stride = p.path_stride;
  cx.epoch = 1;
  cx.path_nodes = 0;
  cx.new_nodes = 0;
  cx.error = 0;
  cx.node_count = 0;
  cx.root_claimed = false;
  const SearchParams sp = p.sp;

  int phase = live ? 0 : 2;  // 0: needs a game, 1: searching, 2: no more games
  unsigned int gi = 0;
  GState g;
  g.cheese = 0; g.p1 = g.p2 = 0; g.mud1 = g.mud2 = 0; g.s1x2 = g.s2x2 = 0;
  int turn = 0;
  Rng rng = rng_seed(0);
  int cheese_available = 0;
  uint32_t n_pos = 0, remaining = 0, nn = 0, term = 0, coll = 0;
  unsigned long long tot_sims = 0, tot_nn = 0, tot_term = 0, tot_coll = 0;

#ifdef AR_HALF_IDLE  // measurement build: cycles a half spends waiting for the other one at the top of the loop
  long long t_top = clock64(), t_end = t_top;
  unsigned long long idle_cycles = 0, total_cycles = 0;
  cx.idle_pick = cx.idle_backup = 0;
#endif
  for (;;) {
    if (__all_sync(FULL, phase == 2)) break;
#ifdef AR_HALF_IDLE
    {
      const long long now = clock64();
      if (phase != 2) { idle_cycles += now - t_end; total_cycles += now - t_top; }
      t_top = now;
    }
#endif
    if (phase == 0) {
      // game_worker_loop (selfplay.rs:609-650): claim the next game index
      unsigned int x = 0;
      if (hl == 0) x = atomicAdd(p.next_game, 1u);
      x = __shfl_sync(__activemask(), x, 0, 16);
      if (x >= (unsigned)p.n_games) {
        phase = 2;
      } else {
        gi = x;
        hw::load_game(p.games + gi, cx, g, turn, hl);
        rng = rng_seed(p.seeds[gi]);
        cheese_available = __popcll(g.cheese);
        hw::init_root(cx, g, hl);
        n_pos = 0;
        tot_sims = tot_nn = tot_term = tot_coll = 0;
        remaining = sp.n_sims;
        nn = term = coll = 0;
        phase = 1;
        if (!p.search_only) {
          uint32_t* co = reinterpret_cast<uint32_t*>(p.summaries[gi].cheese_outcomes);
          for (int i = hl; i < AR_MAX_CELLS / 4; i += 16) co[i] = 0x02020202u;
          if (game_over(g, turn, cx.max_turns)) remaining = 0;  // a game that is over before it starts: summary only
        }
      }
    }
    if (phase == 1) {
      bool finished = !p.search_only && remaining == 0 && n_pos == 0 && game_over(g, turn, cx.max_turns);
      if (!finished) {
        if (remaining > 0) {
          // ---- one simulate_batch of run_search (search.rs:362-390)
          const uint32_t bs = min(remaining, sp.batch_size);
          const uint32_t nn0 = nn, term0 = term;
          hw::simulate_batch_uniform(cx, sp, rng, g, turn, bs, nn, term, coll, p.coll_table_len, hl);
          uint32_t produced = (nn - nn0) + (term - term0);
          produced = produced > 1u ? produced : 1u;
          remaining = remaining > produced ? remaining - produced : 0u;
          if (cx.error) phase = 2;
        }
        if (phase == 1 && remaining == 0) {
          // ---- search finished: extract_result, then one self-play move (selfplay.rs:547-565)
          float pol1[5], pol2[5];
          ar_position_record* pos = p.positions ? p.positions + (size_t)gi * p.pos_stride : nullptr;
          ar_search_result* rout = p.search_only ? (p.search_out + gi) : &pos[n_pos].search;
          uint32_t tv = 0;
          extract_and_store_h(cx, sp, hl, rout, nn, term, coll, pol1, pol2, tv);
          if (p.search_only) {
            phase = 0;
          } else {
            tot_sims += tv; tot_nn += nn; tot_term += term; tot_coll += coll;
            const int a1 = rng_sample_action(rng, pol1);
            const int a2 = rng_sample_action(rng, pol2);
            if (hl == 0) {
              ar_position_record& pr = pos[n_pos];
              pr.p1_x = (uint8_t)(g.p1 % cx.w); pr.p1_y = (uint8_t)(g.p1 / cx.w);
              pr.p2_x = (uint8_t)(g.p2 % cx.w); pr.p2_y = (uint8_t)(g.p2 / cx.w);
              pr.p1_mud = (uint8_t)g.mud1; pr.p2_mud = (uint8_t)g.mud2;
              pr.action_p1 = (uint8_t)a1; pr.action_p2 = (uint8_t)a2;
              pr.turn = (uint16_t)turn; pr.reserved = 0;
              pr.p1_score = 0.5f * (float)g.s1x2; pr.p2_score = 0.5f * (float)g.s2x2;
              uint32_t* cb = reinterpret_cast<uint32_t*>(pr.cheese);
              cb[0] = (uint32_t)g.cheese; cb[1] = (uint32_t)(g.cheese >> 32);
#pragma unroll
              for (int t = 2; t < 8; ++t) cb[t] = 0;
            }
            n_pos += 1;
            // advance_root maps raw actions through action_to_outcome_idx (tree.rs:283-295)
            const uint32_t rmeta = cx.pool[0].s[LANE_LINKS].y;
            const int i = action_to_idx(meta_m1(rmeta), a1), j = action_to_idx(meta_m2(rmeta), a2);
            const uint32_t child = reinterpret_cast<const uint32_t*>(&cx.pool[0].s[LANE_CHILD])[i * 5 + j];
            const uint64_t cheese_before = g.cheese;
            game_step<hw::H_STRIDE>(g, i, j, cx.steptbl());
            turn += 1;
            if (hl == 0) credit_cheese(p.summaries[gi], cheese_before, g);
            hw::hsync();
            remaining = sp.n_sims;
            nn = term = coll = 0;
            if (game_over(g, turn, cx.max_turns)) {
              finished = true;  // the tree of a finished game is dropped
            } else if (child != 0) {
              hw::compact_subtree(cx, child, hl);
            } else {
              hw::init_root(cx, g, hl);  // reinit, tree.rs:298-302
            }
          }
        }
      }
      if (finished) {
        if (hl == 0) {
          ar_game_summary& s = p.summaries[gi];
          s.game_index = gi;
          s.n_positions = n_pos;
          s.final_p1_score = 0.5f * (float)g.s1x2;
          s.final_p2_score = 0.5f * (float)g.s2x2;
          s.result = g.s1x2 > g.s2x2 ? 1 : (g.s2x2 > g.s1x2 ? 2 : 0);
          s.reserved[0] = s.reserved[1] = s.reserved[2] = 0;
          s.cheese_available = (uint16_t)cheese_available;
          s.reserved1 = 0;
          s.total_simulations = tot_sims;
          s.total_nn_evals = tot_nn;
          s.total_terminals = tot_term;
          s.total_collisions = tot_coll;
          if (p.progress) {
            atomicAdd_system((unsigned long long*)&p.progress->positions_completed, (unsigned long long)n_pos);
            atomicAdd_system((unsigned long long*)&p.progress->simulations_completed, tot_sims);
            atomicAdd_system((unsigned long long*)&p.progress->nn_evals_completed, tot_nn);
            atomicAdd_system((unsigned int*)&p.progress->games_completed, 1u);
          }
#ifndef AR_HALF_IDLE
          atomicAdd(&p.counters[2], tot_nn);
          atomicAdd(&p.counters[3], tot_term);
#endif
          atomicAdd(&p.counters[4], (unsigned long long)n_pos);
          atomicAdd(&p.counters[5], tot_sims);
        }
        phase = 0;
      }
    }
#ifdef AR_HALF_IDLE
    t_end = clock64();
#endif
  }
#ifdef AR_HALF_IDLE
  // reported in the path_nodes / new_nodes statistics of this build: waiting cycles / 1024 (end of the descent loop
  // in the low 32 bits... both sums), total cycles / 1024
  if (hl == 0 && live) {
    atomicAdd(&p.counters[2], (unsigned long long)(cx.idle_pick >> 10));
    atomicAdd(&p.counters[3], (unsigned long long)(cx.idle_backup >> 10));
  }
  cx.path_nodes = (uint32_t)(idle_cycles >> 10);
  cx.new_nodes = (uint32_t)(total_cycles >> 10);
#endif
  if (hl == 0 && live) {
    atomicAdd(&p.counters[0], (unsigned long long)cx.path_nodes);
    atomicAdd(&p.counters[1], (unsigned long long)cx.new_nodes);
    if (cx.error) atomicCAS(p.error_flag, 0, (int)cx.error);
  }
  if (p.slot_bitmap) {
    __syncthreads();
    if (threadIdx.x == 0) atomicAnd(&p.slot_bitmap[group >> 5], ~(1u << (group & 31)));
  }
}

// =========================================================================================
// Uniform-prior self-play / search, thread-per-tree (tree_thread.cuh).  One persistent launch:
// every thread claims games from an atomic counter (game_worker_loop, selfplay.rs:609-650) and
// plays them to completion on device (play_game, selfplay.rs:515-598); SmartUniformBackend is
// fused (priors written at node creation, leaf values 0).  The loop condition is warp-uniform so
// that the 32 trees of a warp reconverge at every step.
// =========================================================================================
constexpr int TT_BLOCK = 32;
template <int MIN_BLOCKS, int NW>
__global__ void __launch_bounds__(TT_BLOCK, MIN_BLOCKS) selfplay_tt_kernel(tt::Ctx c, int n_slots) {
  using T = tt::TT<NW>;
  extern __shared__ __align__(16) uint32_t tt_smem[];  // maze image: 16 * NW words per tree, word-interleaved
  const int slot = blockIdx.x * TT_BLOCK + threadIdx.x;
  typename T::TState s;
  typename T::TArr a;
  T::tt_init(s, c, (uint32_t)slot, tt_smem + threadIdx.x, TT_BLOCK);
  if (slot >= n_slots) s.phase = tt::PH_EXIT;
  const int lane = threadIdx.x & 31;
  // Warp-level phase scheduler: every iteration runs ONE phase, the one most of the warp's trees are in;
  // the others wait (they cost nothing) and pile up until their phase is the majority.  A descend step is
  // ~8x the instructions of a backup step, so running both every iteration (plain divergence) left 5.7 of 32
  // lanes active (profiles/r2_summary.md section 3).
  for (;;) {
    const unsigned md = __ballot_sync(0xffffffffu, s.phase == tt::PH_DESCEND);
    const unsigned mb = __ballot_sync(0xffffffffu, s.phase == tt::PH_BACKUP);
    const unsigned mc = __ballot_sync(0xffffffffu, s.phase == tt::PH_CONTROL);
    if (!(md | mb | mc)) break;
    if (mc) {  // rare (a few per move) and it unblocks a tree: always first
      const unsigned need = __ballot_sync(0xffffffffu, s.phase == tt::PH_CONTROL && s.cstate == tt::CS_COMPACT_MARK);
      if (need) {
        T::coop_compact(s, c, need, lane);
      } else if (s.phase == tt::PH_CONTROL) {
        T::template control<true>(s, c);
      }
      continue;
    }
    if (__popc(mb) >= __popc(md)) {
      if (s.phase == tt::PH_BACKUP) T::template step_backup<true>(s, a, c);
    } else {
      if (s.phase == tt::PH_DESCEND) T::template step_descend<true>(s, a, c);
    }
  }
  atomicAdd(&c.counters[0], s.path_nodes);
  atomicAdd(&c.counters[1], s.new_nodes);
  if (s.error) atomicCAS(c.error_flag, 0, (int)s.error);
}

// Block-sorted variant: the trees' state lives in shared memory instead of registers, and every
// iteration the block sorts its trees by phase, so a warp runs 32 trees that are all in the SAME phase
// (descend / backup / control) — only the warps at a phase boundary diverge.  Registers are bound to
// work items for one step only; a tree is picked up by whichever lane its rank in the sorted order says.
// (With trees bound to lanes 7.7 of 32 lanes were active on the real workload, profiles/r2_tt_lanes.md.)
constexpr int TB_BLOCK = 128;
constexpr int TB_BACKUP_NODES = 4;
template <int NW>
__host__ __device__ inline size_t tb_smem_bytes() {
  return (size_t)TB_BLOCK * (sizeof(typename tt::TT<NW>::TState) + 4 * tt::TT<NW>::MAZE_WORDS + 2 + 2) + 256;
}
template <int MIN_BLOCKS, int NW>
__global__ void __launch_bounds__(TB_BLOCK, MIN_BLOCKS) selfplay_tb_kernel(tt::Ctx c, typename tt::TT<NW>::TArr* arrs) {
  using T = tt::TT<NW>;
  using TState = typename T::TState;
  using TArr = typename T::TArr;
  extern __shared__ __align__(16) uint8_t tb_smem[];
  TState* st = reinterpret_cast<TState*>(tb_smem);
  uint32_t* maze = reinterpret_cast<uint32_t*>(st + TB_BLOCK);              // [MAZE_WORDS][TB_BLOCK]
  uint16_t* order = reinterpret_cast<uint16_t*>(maze + T::MAZE_WORDS * TB_BLOCK);  // trees sorted by phase
  uint16_t* creq = order + TB_BLOCK;                                        // trees that asked for a compaction
  int* wcnt = reinterpret_cast<int*>(creq + TB_BLOCK);                      // [4 classes][4 warps]
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int NWARP = TB_BLOCK / 32;
  {
    TState s;
    T::tt_init(s, c, (uint32_t)(blockIdx.x * TB_BLOCK + tid), maze + tid, TB_BLOCK);
    st[tid] = s;
  }
  TArr* my_arrs = arrs + (size_t)blockIdx.x * TB_BLOCK;
  __syncthreads();
  for (;;) {
    // ---- classify this thread's own tree: 0 descend, 1 backup, 2 control, 3 compaction request, 4 exited
    const int ph = st[tid].phase;
    int cls = ph == tt::PH_DESCEND ? 0 : ph == tt::PH_BACKUP ? 1 : ph == tt::PH_CONTROL ? 2 : 4;
    if (cls == 2 && st[tid].cstate == tt::CS_COMPACT_MARK) cls = 3;
    unsigned bal[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) bal[k] = __ballot_sync(0xffffffffu, cls == k);
    if (lane < 4) wcnt[lane * NWARP + wid] = __popc(bal[lane]);
    __syncthreads();
    // ---- ranks: class-major, then warp, then lane
    int base[4], total = 0, n_creq = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int before = 0, all = 0;
#pragma unroll
      for (int w = 0; w < NWARP; ++w) {
        const int v = wcnt[k * NWARP + w];
        before += w < wid ? v : 0;
        all += v;
      }
      if (k < 3) { base[k] = total + before; total += all; } else { base[k] = before; n_creq = all; }
    }
    if (total == 0 && n_creq == 0) break;
    if (cls < 3) order[base[cls] + __popc(bal[cls] & ((1u << lane) - 1u))] = (uint16_t)tid;
    else if (cls == 3) creq[base[3] + __popc(bal[3] & ((1u << lane) - 1u))] = (uint16_t)tid;
    __syncthreads();
    // ---- compaction requests: one warp per tree, all 32 lanes (tree_thread.cuh coop_compact_tree)
    for (int r = wid; r < n_creq; r += NWARP) {
      TState& ts = st[creq[r]];
      const uint32_t kept = T::coop_compact_tree(ts.pt, c.arena, ts.cp_count, ts.cp_new_root, ts.cp_page0, ts.cp_page1,
                                                  ts.cp_page2, ts.cp_page3, lane);
      if (lane == 0) {
        ts.cp_kept = kept;
        T::compact_end(ts, c);
        ts.cstate = tt::CS_MOVE_START;
      }
      __syncwarp();
    }
    // ---- one step for the tree of this thread's rank
    if (tid < total) {
      const int tree = order[tid];
      TState s = st[tree];
      TArr& a = my_arrs[tree];
      if (s.phase == tt::PH_DESCEND) T::template step_descend<true>(s, a, c);
      else if (s.phase == tt::PH_BACKUP) T::template step_backup<true>(s, a, c, TB_BACKUP_NODES);
      else T::template control<true>(s, c);
      st[tree] = s;
    }
    __syncthreads();
  }
  atomicAdd(&c.counters[0], st[tid].path_nodes);
  atomicAdd(&c.counters[1], st[tid].new_nodes);
  if (st[tid].error) atomicCAS(c.error_flag, 0, (int)st[tid].error);
}

// =========================================================================================
// NN-guided mode: the search is cut at the evaluator.  One step = tree kernel (back up the
// previous batch with the evaluator's outputs, finish moves / games, gather the next batch and
// append its leaves to the global evaluation queue) followed by the fused encode+MLP kernel
// over the queue (nn_kernels.cu).  The host only enqueues steps; it reads one counter every few
// steps to know when every game is finished.  Replaces MuxBackend's cross-game batching
// (crates/alpharat-sampling/src/backends/mux.rs:170-289).
// =========================================================================================
enum SlotPhase : uint32_t { PH_IDLE = 0, PH_GATHER = 1, PH_WAIT_EVAL = 2, PH_DONE = 3, PH_COMPACT = 4 };
// Loop iterations of the tree compaction (32 nodes marked or 4 records moved each) a slot may spend
// in one step: about the cost of gathering one batch.
constexpr int COMPACT_ITERS_PER_STEP = 48;

struct SlotState {
  GPack g;
  Rng rng;
  unsigned long long tot_sims, tot_nn, tot_term, tot_coll;
  int turn, gi;
  uint32_t node_count, epoch;
  uint32_t remaining, nn, term, coll;
  uint32_t n_pos, cheese_available;
  uint32_t n_tp, row_base;
  uint32_t phase, error;
  uint32_t path_nodes, new_nodes;
  CompactState comp;
  uint32_t pad[1];
};

// Evaluation cache (CachedBackend / NNCache, crates/alpharat-sampling/src/cached_backend.rs:54-120,
// nn_cache.rs): one direct-mapped table per resident tree, the GPU counterpart of the reference's
// thread-local caches.  A leaf whose position was already scored skips the evaluator; the search cannot
// tell (it still counts as an nn_eval), so results are bit-identical with and without the cache.  The key
// is the whole position (not only a hash), tagged with the game index and the turn.
struct __align__(16) CacheEnt {
  uint64_t cheese;
  uint32_t pos, score;      // GPack
  uint32_t game_idx;
  uint32_t turn_tag;        // turn | 0x80000000 (0 = empty)
  uint32_t pad[2];
  float out[12];
};
static_assert(sizeof(CacheEnt) == 80, "CacheEnt layout");
__device__ __forceinline__ uint32_t cache_index(const GPack& g, uint32_t game, uint32_t turn_tag, uint32_t mask) {
  uint64_t h = g.cheese * 0x9E3779B97F4A7C15ull;
  h ^= (((uint64_t)g.pos << 32) | g.score) * 0xC2B2AE3D27D4EB4Full;
  h ^= (((uint64_t)game << 20) ^ turn_tag) * 0x165667B19E3779F9ull;
  h ^= h >> 29;
  h *= 0xBF58476D1CE4E5B9ull;
  h ^= h >> 32;
  return (uint32_t)h & mask;
}

struct NnParams {
  SlotState* slots;
  TpEntry* tp_store;        // [n_slots][batch_cap]
  EvalRow* rows;            // evaluation queue
  uint32_t* n_rows;         // queue length (reset by the host every step)
  const float* nn_out;      // [rows][12] from the previous step
  uint32_t* done_slots;
  uint32_t max_rows;
  CacheEnt* cache;          // [n_slots][cache_mask + 1] or null
  GPack* key_store;         // [n_slots][batch_cap]: positions of the leaves waiting for their evaluation
  uint32_t cache_mask;
  int max_iters;            // gather / backup rounds a slot may chain inside one step
  uint4* board_store;       // [n_slots][SM_TP / 16]: move table + maze of the slot's game, built once per game
  int slot_begin, slot_end; // the slots this launch steps (the slots are split into groups that alternate)
};

__global__ void nn_init_slots_kernel(SlotState* slots, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  SlotState st;
  memset(&st, 0, sizeof(st));
  st.phase = PH_IDLE;
  st.gi = -1;
  st.epoch = 1;
  slots[i] = st;
}

// WPB = warps per block (AR_NN_WPB, default 1): the step is bulk-synchronous and its slots finish at very different
// times (a 16-leaf batch, a one-leaf batch, a slice of a tree compaction); with one warp per block a finished slot
// frees its share of the SM at once for the other slot group's step kernel.
template <int WPB>
__global__ void __launch_bounds__(32 * WPB, AR_MIN_BLOCKS * 4 / WPB) nn_step_kernel(RunParams p, NnParams q) {
  extern __shared__ __align__(16) uint8_t smem[];
  int lane = threadIdx.x & 31;
  asm volatile("" : "+r"(lane));
  const int wib = threadIdx.x >> 5;
  const int slot = q.slot_begin + blockIdx.x * WPB + wib;
  if (slot >= q.slot_end) return;
  SlotState* sp_g = q.slots + slot;
  if (sp_g->phase == PH_DONE) return;

  uint8_t* base = smem + (size_t)wib * warp_smem_bytes(p.max_depth, p.batch_cap);
  WarpCtx cx;
  cx.bind(base, p.pools + (size_t)slot * p.pool_nodes, lane, p.max_depth, p.batch_cap);
  cx.path_buf = p.path_bufs + (size_t)slot * p.batch_cap * p.path_stride;
  cx.remap = p.remaps + (size_t)slot * p.pool_nodes;
  cx.coll_table = p.coll_table;
  cx.pool_nodes = p.pool_nodes;
  cx.path_stride = p.path_stride;
  const SearchParams sp = p.sp;

  SlotState st = *sp_g;
  cx.epoch = st.epoch;
  cx.node_count = st.node_count;
  cx.path_nodes = st.path_nodes;
  cx.new_nodes = st.new_nodes;
  cx.error = st.error;
  cx.root_claimed = false;
  GState g = g_unpack(st.g);
  Rng rng = st.rng;
  int turn = st.turn;
  uint4* const board_g = q.board_store + (size_t)slot * (SM_TP / 16);
  if (st.gi >= 0) {
    // the game's move table + maze (SM_TP bytes at the start of the warp's shared memory) were built when the
    // game was loaded: copy them back instead of rebuilding them every step
    const ar_game_pod* pod = p.games + st.gi;
    cx.w = pod->width;
    cx.cells = (int)pod->width * pod->height;
    cx.max_turns = pod->max_turns;
    for (int i = lane; i < SM_TP / 16; i += 32) reinterpret_cast<uint4*>(cx.sm)[i] = board_g[i];
    __syncwarp();
  }
  TpEntry* tp_g = q.tp_store + (size_t)slot * p.batch_cap;

  CacheEnt* const cbase = q.cache ? q.cache + (size_t)slot * (q.cache_mask + 1) : nullptr;
  GPack* const keys_g = q.key_store + (size_t)slot * p.batch_cap;
  // populate + backup in to_process order (search.rs:1028-1058); the batch entries are in cx.tp().
  // kind 0: scored by the evaluator in the previous step, 1: terminal, 2: evaluation-cache hit.
  auto process_batch = [&](int n_tp) {
    uint32_t row = st.row_base, nn_b = 0, term_b = 0;
    for (int e = 0; e < n_tp; ++e) {
      const TpEntry te = cx.tp()[e];
      if (te.kind == 1) {
        term_b += 1;
        backup_entry(cx, e, 0.0f, 0.0f, nullptr, nullptr, lane);
        continue;
      }
      const float* o = te.kind == 0 ? q.nn_out + (size_t)row * 12 : cbase[te.pad].out;
#ifdef AR_CACHE_VERIFY  // debug build: hits are evaluated anyway and compared with the cached outputs
      if (te.kind == 0 && te.pad != 0 && lane < 12)
        atomicAdd(&p.counters[cbase[te.pad - 1].out[lane] != o[lane] ? 5 : 4], 1ull);
#endif
      row += te.kind == 0 ? 1u : 0u;
      nn_b += 1;
      if (te.node == 0 && sp.noise_epsilon > 0.0f) {
        // populate first, then noise, then backup (search.rs:1034-1052)
        backup_entry(cx, e, o[10], o[11], o, o + 5, lane, /*populate_only=*/true);
        apply_root_noise(cx, sp, rng, lane);
        backup_entry(cx, e, o[10], o[11], nullptr, nullptr, lane);
      } else {
        backup_entry(cx, e, o[10], o[11], o, o + 5, lane);
      }
    }
    if (cbase) {  // remember what the evaluator returned (after every hit of this batch has been read)
      uint32_t rows_before = 0;
      for (int e0 = 0; e0 < n_tp; e0 += 32) {
        const int e = e0 + lane;
        const bool fresh = e < n_tp && cx.tp()[e].kind == 0;
        const uint32_t fm = __ballot_sync(FULL, fresh);
        if (fresh) {
          const uint32_t r = st.row_base + rows_before + __popc(fm & ((1u << lane) - 1u));
          const GPack gp = keys_g[e];
          const uint32_t tt = (uint32_t)(turn + cx.tp()[e].depth) | 0x80000000u;
          CacheEnt* ce = cbase + cache_index(gp, (uint32_t)st.gi, tt, q.cache_mask);
          const float4* src = reinterpret_cast<const float4*>(q.nn_out + (size_t)r * 12);
          *reinterpret_cast<uint4*>(ce) = make_uint4((uint32_t)gp.cheese, (uint32_t)(gp.cheese >> 32), gp.pos, gp.score);
          *reinterpret_cast<uint4*>(&ce->game_idx) = make_uint4((uint32_t)st.gi, tt, 0u, 0u);
          float4* dst = reinterpret_cast<float4*>(ce->out);
          dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
        }
        rows_before += __popc(fm);
      }
    }
    st.nn += nn_b;
    st.term += term_b;
    uint32_t produced = nn_b + term_b;
    produced = produced > 1u ? produced : 1u;
    st.remaining = st.remaining > produced ? st.remaining - produced : 0u;
  };

  bool moved = false;
  for (int iter = 0; iter < q.max_iters && cx.error == 0; ++iter) {
    if (st.phase == PH_IDLE) {
      unsigned int gi = 0;
      if (lane == 0) gi = atomicAdd(p.next_game, 1u);
      gi = __shfl_sync(FULL, gi, 0);
      if (gi >= (unsigned)p.n_games) {
        st.phase = PH_DONE;
        st.gi = -1;
        if (lane == 0) atomicAdd(q.done_slots, 1u);
        break;
      }
      st.gi = (int)gi;
      load_game(p.games + gi, cx, g, turn, lane);
      for (int i = lane; i < SM_TP / 16; i += 32) board_g[i] = reinterpret_cast<const uint4*>(cx.sm)[i];
      rng = rng_seed(p.seeds[gi]);
      st.cheese_available = __popcll(g.cheese);
      cx.epoch += 1;
      init_root(cx, g, lane);
      if (!p.search_only) init_cheese_outcomes(p.summaries[gi], lane);
      st.n_pos = 0;
      st.tot_sims = st.tot_nn = st.tot_term = st.tot_coll = 0;
      st.remaining = sp.n_sims;
      st.nn = st.term = st.coll = 0;
      st.phase = PH_GATHER;
      if (!p.search_only && game_over(g, turn, cx.max_turns)) st.remaining = 0;  // empty game
    }
    if (st.phase == PH_COMPACT) {
      int budget = COMPACT_ITERS_PER_STEP;
      if (!compact_step(cx, st.comp, lane, budget)) break;
      st.phase = PH_GATHER;
      if (budget < COMPACT_ITERS_PER_STEP / 2) break;  // gather in the next step
    }
    if (st.phase == PH_WAIT_EVAL) {
      const int n_tp = (int)st.n_tp;
      for (int e = lane; e < n_tp; e += 32) cx.tp()[e] = tp_g[e];
      __syncwarp();
      process_batch(n_tp);
      st.phase = PH_GATHER;
    }
    if (st.phase == PH_GATHER) {
      if (st.remaining == 0) {
        // ---- search finished: extract_result, then one self-play move (selfplay.rs:547-565)
        const bool empty_game = !p.search_only && game_over(g, turn, cx.max_turns);
        if (!empty_game) {
          float pol1[5], pol2[5];
          ar_position_record* pos = p.positions ? p.positions + (size_t)st.gi * p.pos_stride : nullptr;
          ar_search_result* rout = p.search_only ? (p.search_out + st.gi) : &pos[st.n_pos].search;
          extract_and_store(cx, sp, lane, rout, st.nn, st.term, st.coll, pol1, pol2);
          if (p.search_only) {
            st.phase = PH_IDLE;
            continue;
          }
          uint32_t tv = 0;
          if (lane == 0) tv = rout->total_visits;
          tv = __shfl_sync(FULL, tv, 0);
          st.tot_sims += tv; st.tot_nn += st.nn; st.tot_term += st.term; st.tot_coll += st.coll;
          int a1 = rng_sample_action(rng, pol1);
          int a2 = rng_sample_action(rng, pol2);
          if (lane == 0) {
            ar_position_record& pr = pos[st.n_pos];
            pr.p1_x = (uint8_t)(g.p1 % cx.w); pr.p1_y = (uint8_t)(g.p1 / cx.w);
            pr.p2_x = (uint8_t)(g.p2 % cx.w); pr.p2_y = (uint8_t)(g.p2 / cx.w);
            pr.p1_mud = (uint8_t)g.mud1; pr.p2_mud = (uint8_t)g.mud2;
            pr.action_p1 = (uint8_t)a1; pr.action_p2 = (uint8_t)a2;
            pr.turn = (uint16_t)turn; pr.reserved = 0;
            pr.p1_score = 0.5f * (float)g.s1x2; pr.p2_score = 0.5f * (float)g.s2x2;
            uint32_t* cb = reinterpret_cast<uint32_t*>(pr.cheese);
            cb[0] = (uint32_t)g.cheese; cb[1] = (uint32_t)(g.cheese >> 32);
#pragma unroll
            for (int t = 2; t < 8; ++t) cb[t] = 0;
          }
          st.n_pos += 1;
          uint32_t rmeta = cx.pool[0].s[LANE_LINKS].y;
          int i = action_to_idx(meta_m1(rmeta), a1), j = action_to_idx(meta_m2(rmeta), a2);
          uint32_t child = reinterpret_cast<const uint32_t*>(&cx.pool[0].s[LANE_CHILD])[i * 5 + j];
          const uint64_t cheese_before = g.cheese;
          game_step(g, i, j, cx.steptbl());  // indexed by outcome
          turn += 1;
          if (lane == 0) credit_cheese(p.summaries[st.gi], cheese_before, g);
          __syncwarp();
          moved = true;
          if (!game_over(g, turn, cx.max_turns)) {  // the tree of a finished game is dropped
            if (child != 0) {
              st.comp = compact_begin(cx, child);  // advance_root, spread over the next steps
              st.phase = PH_COMPACT;
            } else {
              cx.epoch += 1;
              init_root(cx, g, lane);  // reinit, tree.rs:298-302
            }
          }
        }
        if (game_over(g, turn, cx.max_turns)) {
          if (lane == 0) {
            ar_game_summary& s = p.summaries[st.gi];
            s.game_index = (uint32_t)st.gi;
            s.n_positions = st.n_pos;
            s.final_p1_score = 0.5f * (float)g.s1x2;
            s.final_p2_score = 0.5f * (float)g.s2x2;
            s.result = g.s1x2 > g.s2x2 ? 1 : (g.s2x2 > g.s1x2 ? 2 : 0);
            s.cheese_available = (uint16_t)st.cheese_available;
            s.total_simulations = st.tot_sims;
            s.total_nn_evals = st.tot_nn;
            s.total_terminals = st.tot_term;
            s.total_collisions = st.tot_coll;
            s.reserved[0] = s.reserved[1] = s.reserved[2] = 0;
            s.reserved1 = 0;
            if (p.progress) {
              atomicAdd_system((unsigned long long*)&p.progress->positions_completed, (unsigned long long)st.n_pos);
              atomicAdd_system((unsigned long long*)&p.progress->simulations_completed, st.tot_sims);
              atomicAdd_system((unsigned long long*)&p.progress->nn_evals_completed, st.tot_nn);
              atomicAdd_system((unsigned int*)&p.progress->games_completed, 1u);
            }
          }
          st.phase = PH_IDLE;
          continue;
        }
        st.remaining = sp.n_sims;
        st.nn = st.term = st.coll = 0;
        // the step is bulk-synchronous: a slot that just moved starts its next search (or the
        // compaction of its tree) in the next step, so that the step's critical path stays short
        if (moved) break;
      }
      // ---- gather one batch (simulate_batch up to the evaluator, search.rs:961-1023)
      cx.epoch += 1;
      cx.root_claimed = false;
      const uint32_t bs = min(st.remaining, sp.batch_size);
      uint32_t ci = cx.node_count < p.coll_table_len ? cx.node_count : p.coll_table_len - 1;
      int collisions_left = (int)cx.coll_table[ci];
      int n_tp = 0;
      while ((uint32_t)n_tp < bs && collisions_left > 0 && cx.error == 0) {
        uint32_t budget = min((uint32_t)collisions_left, bs - (uint32_t)n_tp);
        uint32_t c = pick_nodes<true>(cx, sp, rng, g, turn, budget, n_tp, false, lane);
        collisions_left -= (int)c;
        st.coll += c;
      }
      if (cx.error) break;
      __syncwarp();
      uint32_t n_hit = 0;
      if (cbase) {  // evaluation-cache lookup: hits become kind 2 and never reach the queue
        for (int e0 = 0; e0 < n_tp; e0 += 32) {
          const int e = e0 + lane;
          bool hit = false;
          if (e < n_tp) {
            TpEntry te = cx.tp()[e];
            if (te.kind == 0) {
              const GPack gp = cx.tp_state()[e];
              const uint32_t tt = (uint32_t)(turn + te.depth) | 0x80000000u;
              const uint32_t idx = cache_index(gp, (uint32_t)st.gi, tt, q.cache_mask);
              const uint4 k0 = *reinterpret_cast<const uint4*>(cbase + idx);
              const uint2 k1 = *reinterpret_cast<const uint2*>(&cbase[idx].game_idx);
              hit = k0.x == (uint32_t)gp.cheese && k0.y == (uint32_t)(gp.cheese >> 32) && k0.z == gp.pos &&
                    k0.w == gp.score && k1.x == (uint32_t)st.gi && k1.y == tt;
              if (hit) {
#ifdef AR_CACHE_VERIFY
                te.pad = (uint16_t)(idx + 1);
                cx.tp()[e] = te;
                keys_g[e] = gp;
                hit = false;
#else
                te.kind = 2;
                te.pad = (uint16_t)idx;
                cx.tp()[e] = te;
#endif
              } else {
                keys_g[e] = gp;
              }
            }
          }
          n_hit += __popc(__ballot_sync(FULL, hit));
        }
        __syncwarp();
      }
      uint32_t evmask = __ballot_sync(FULL, lane < n_tp && cx.tp()[lane].kind == 0);
      uint32_t evmask_hi = n_tp > 32 ? __ballot_sync(FULL, lane + 32 < n_tp && cx.tp()[lane + 32].kind == 0) : 0u;
      const uint32_t n_eval = __popc(evmask) + __popc(evmask_hi);
      st.n_tp = (uint32_t)n_tp;
      if (cbase && lane == 0 && (n_hit | n_eval)) {
        atomicAdd(&p.counters[6], (unsigned long long)n_hit);
        atomicAdd(&p.counters[7], (unsigned long long)n_eval);
      }
      if (n_eval == 0) {
        // nothing to evaluate (terminals and cache hits only): back up at once and keep going
        process_batch(n_tp);
        continue;
      }
      uint32_t rb = 0;
      if (lane == 0) rb = atomicAdd(q.n_rows, n_eval);
      rb = __shfl_sync(FULL, rb, 0);
      if (rb + n_eval > q.max_rows) { cx.error = AR_ERR_POOL_OVERFLOW; break; }
      st.row_base = rb;
      for (int e = lane; e < n_tp; e += 32) {
        TpEntry te = cx.tp()[e];
        tp_g[e] = te;
        if (te.kind == 0) {
          uint32_t m = e < 32 ? (evmask & ((1u << e) - 1u)) : evmask;
          uint32_t mh = e < 32 ? 0u : (evmask_hi & ((1u << (e - 32)) - 1u));
          uint32_t r = rb + __popc(m) + __popc(mh);
          GPack gp = cx.tp_state()[e];
          EvalRow er;
          er.cheese = gp.cheese;
          er.pos = gp.pos;
          er.score = gp.score;
          er.game_idx = (uint32_t)st.gi;
          er.turn = (uint16_t)(turn + te.depth);
          er.max_turns = (uint16_t)cx.max_turns;
          er.pad[0] = er.pad[1] = 0;
          q.rows[r] = er;
        }
      }
      st.phase = PH_WAIT_EVAL;
      break;
    }
  }

  st.g = g_pack(g);
  st.rng = rng;
  st.turn = turn;
  st.epoch = cx.epoch;
  st.node_count = cx.node_count;
  st.path_nodes = cx.path_nodes;
  st.new_nodes = cx.new_nodes;
  st.error = cx.error;
  __syncwarp();
  if (lane == 0) {
    *sp_g = st;
    if (st.phase == PH_DONE) {
      atomicAdd(&p.counters[0], (unsigned long long)cx.path_nodes);
      atomicAdd(&p.counters[1], (unsigned long long)cx.new_nodes);
    }
    if (cx.error) atomicCAS(p.error_flag, 0, (int)cx.error);
  }
}

// Games end long before max_turns on average: pack the records that exist (one warp per game)
// so that the device-to-host copy moves only those.
__global__ void compact_positions_kernel(const ar_position_record* __restrict__ src, int stride,
                                         const uint32_t* __restrict__ offsets, int n,
                                         ar_position_record* __restrict__ dst) {
  const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (g >= n) return;
  const uint32_t b = offsets[g], e = offsets[g + 1];
  const uint32_t* s = reinterpret_cast<const uint32_t*>(src + (size_t)g * stride);
  uint32_t* d = reinterpret_cast<uint32_t*>(dst + b);
  const uint32_t words = (e - b) * (uint32_t)(sizeof(ar_position_record) / 4);
  for (uint32_t i = lane; i < words; i += 32) d[i] = s[i];
}

}  // namespace ar

// =========================================================================================
// Host side: engine object and C-ABI
// =========================================================================================
using namespace ar;

// Everything one batch of games needs on the device and on its way back: inputs, outputs, the work
// counter the persistent kernel claims games from, a stream and two events.  The engine's blocking API
// uses `main`; the streaming API (ar_stream_*) keeps several in flight.
struct BatchBuf {
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  ar_game_pod* d_games = nullptr;
  uint64_t* d_seeds = nullptr;
  ar_game_summary* d_summaries = nullptr;
  ar_position_record* d_positions = nullptr;
  int cap_games = 0, cap_stride = 0;
  ar_position_record* d_dense = nullptr;   // packed records for the download
  ar_position_record* h_dense = nullptr;   // pinned staging
  uint32_t* d_offsets = nullptr;
  size_t cap_dense = 0;
  int cap_offsets = 0;
  int n_resident = 0, resident_stride = 0;
  bool resident_valid = false;  // d_games / d_seeds hold an uploaded batch (search and evaluator calls overwrite them)
  std::vector<ar_game_pod> h_games;  // board sizes of the resident batch (evaluator shape check)
  unsigned int* d_next = nullptr;
  unsigned long long* d_counters = nullptr;
  int* d_error = nullptr;
  uint64_t h2d = 0, d2h = 0, launches = 0;
  // streaming only
  ar_game_pod* h_games_pinned = nullptr;
  uint64_t* h_seeds_pinned = nullptr;
  bool in_flight = false;
  std::chrono::steady_clock::time_point t_submit;
};

struct ar_engine {
  ar_engine_cfg cfg{};
  int device = 0;
  BatchBuf main;
  BatchBuf* cur = &main;                       // the batch the internal helpers operate on
  std::vector<BatchBuf*> sbufs;                // ar_stream_open
  uint32_t* d_slot_bitmap = nullptr;           // tree slots in use by streaming launches
  cudaEvent_t ev_base = nullptr;               // time origin of ar_stream_times
  cudaStream_t stream2 = nullptr;              // second slot group of the NN-guided loop
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  std::string err;
  // thread-per-tree uniform engine (tree_thread.cuh): paged node arena shared by all resident trees
  uint8_t* tt_arena = nullptr;
  uint32_t* tt_bitmap = nullptr;
  uint32_t* tt_page_tables = nullptr;
  uint32_t tt_n_pages = 0, tt_bitmap_words = 0, tt_pt_stride = 0, tt_slots = 0;
  void* tt_arrs = nullptr;      // per-tree cell stack + batch entries of the block-sorted kernel (TT<NW>::TArr)
  // per-slot storage of the warp-per-tree NN-guided engine (allocated on first use)
  NodeRec* pools = nullptr;
  uint32_t* path_bufs = nullptr;
  uint32_t* remaps = nullptr;
  uint16_t* coll_table = nullptr;
  uint32_t pool_nodes = 0, path_stride = 0, max_depth = 0, batch_cap = 0, n_slots = 0;
  uint32_t node_cap = 0;  // nodes a tree may hold (pool_nodes of the engine cfg, or the worst case)
  ar_search_cfg coll_cfg{};
  uint32_t coll_len = 0;
  bool coll_valid = false;
  ar_search_result* d_search = nullptr;
  int cap_search = 0;
  ar_progress* h_progress = nullptr;  // mapped pinned
  ar_progress* d_progress = nullptr;
  // leaf evaluator
  int arch = AR_ARCH_UNIFORM;
  int nn_width = 0, nn_height = 0;
  LeafEvaluator* eval = nullptr;
  EvalRow* d_rows = nullptr;
  float* d_nn_out = nullptr;
  int cap_rows = 0;
  uint16_t* d_maze_tab = nullptr;  // per-game bf16 maze channels (build_maze_table)
  int cap_maze = 0;
  // NN-mode step state
  SlotState* d_slots = nullptr;
  TpEntry* d_tp_store = nullptr;
  EvalRow* d_queue = nullptr;
  float* d_queue_out = nullptr;
  uint32_t* d_n_rows = nullptr;   // [0] queue length, [1] done slots
  uint64_t nn_steps = 0;
  // evaluation cache (ar_engine_set_eval_cache)
  CacheEnt* d_cache = nullptr;
  GPack* d_key_store = nullptr;
  uint4* d_board_store = nullptr; // per-slot move table + maze image
  uint32_t cache_entries = 0;     // per resident tree, power of two (0 = disabled)
};

static thread_local std::string g_create_error;

#define CK(call)                                                                          \
  do {                                                                                    \
    cudaError_t _e = (call);                                                              \
    if (_e != cudaSuccess) {                                                              \
      e->err = std::string(#call) + ": " + cudaGetErrorString(_e);                        \
      return AR_ERR_CUDA;                                                                 \
    }                                                                                     \
  } while (0)

static ar_status ensure_coll_table(ar_engine* e, const ar_search_cfg& c) {
  if (e->coll_valid && memcmp(&e->coll_cfg.collision_limit_min, &c.collision_limit_min,
                              sizeof(uint32_t) * 4 + sizeof(float)) == 0)
    return AR_OK;
  // beyond collision_scaling_end the budget is constant, so the table may be clamped
  if (c.collision_scaling_end >= e->coll_len && e->node_cap + 1 > e->coll_len) {
    e->err = "collision_scaling_end beyond the collision table";
    return AR_ERR_UNSUPPORTED;
  }
  std::vector<uint16_t> t = ar_host::collision_table(c, e->coll_len);
  CK(cudaMemcpyAsync(e->coll_table, t.data(), t.size() * sizeof(uint16_t), cudaMemcpyHostToDevice,
                     e->cur->stream));
  CK(cudaStreamSynchronize(e->cur->stream));
  e->coll_cfg = c;
  e->coll_valid = true;
  return AR_OK;
}

static ar_status validate_cfg(ar_engine* e, const ar_search_cfg* c) {
  if (!c) { e->err = "search cfg is NULL"; return AR_ERR_INVALID_ARG; }
  if (c->simulations == 0) { e->err = "simulations must be > 0"; return AR_ERR_INVALID_ARG; }
  if (c->batch_size == 0 || c->batch_size > e->batch_cap) {
    e->err = "batch_size " + std::to_string(c->batch_size) + " outside [1, engine max_batch_size=" +
             std::to_string(e->batch_cap) + "]";
    return AR_ERR_INVALID_ARG;
  }
  if (c->collision_limit_max + 2 * c->batch_size > 1023u) {
    e->err = "collision_limit_max + 2 * batch_size must be <= 1023 (10-bit in-flight counters)";
    return AR_ERR_INVALID_ARG;
  }
  if ((uint64_t)c->simulations * e->cfg.max_turns >= (1ull << VIS_BITS)) {
    e->err = "simulations * max_turns must be < 2^22 (22-bit edge visit counters)";
    return AR_ERR_INVALID_ARG;
  }
  return AR_OK;
}

static ar_status validate_games(ar_engine* e, const ar_game_pod* games, int n) {
  if (n < 0 || (n > 0 && !games)) { e->err = "games is NULL"; return AR_ERR_INVALID_ARG; }
  for (int i = 0; i < n; ++i) {
    const ar_game_pod& g = games[i];
    uint32_t cells = (uint32_t)g.width * g.height;
    if (g.width == 0 || g.height == 0) { e->err = "game " + std::to_string(i) + ": empty board"; return AR_ERR_INVALID_ARG; }
    if (cells > e->cfg.max_cells || g.width > 16 || g.height > 16) {
      e->err = "game " + std::to_string(i) + ": " + std::to_string(g.width) + "x" + std::to_string(g.height) +
               " board exceeds the engine's max_cells (" + std::to_string(e->cfg.max_cells) +
               "; boards over 64 cells need tree_engine = AR_TREE_THREAD and no evaluator)";
      return AR_ERR_UNSUPPORTED;
    }
    if (g.max_turns > e->cfg.max_turns) {
      e->err = "game " + std::to_string(i) + ": max_turns " + std::to_string(g.max_turns) +
               " exceeds engine max_turns " + std::to_string(e->cfg.max_turns);
      return AR_ERR_INVALID_ARG;
    }
    if (g.p1_x >= g.width || g.p2_x >= g.width || g.p1_y >= g.height || g.p2_y >= g.height) {
      e->err = "game " + std::to_string(i) + ": player outside the board";
      return AR_ERR_INVALID_ARG;
    }
  }
  return AR_OK;
}

static void pod_to_row(const ar_game_pod& g, uint32_t game_idx, EvalRow& r) {
  memset(&r, 0, sizeof(r));
  memcpy(&r.cheese, g.cheese, 8);
  uint32_t p1 = (uint32_t)g.p1_y * g.width + g.p1_x, p2 = (uint32_t)g.p2_y * g.width + g.p2_x;
  r.pos = p1 | (p2 << 8) | ((uint32_t)g.p1_mud << 16) | ((uint32_t)g.p2_mud << 24);
  r.score = (uint32_t)lrintf(g.p1_score * 2.0f) | ((uint32_t)lrintf(g.p2_score * 2.0f) << 16);
  r.game_idx = game_idx;
  r.turn = g.turn;
  r.max_turns = g.max_turns;
}

// Stage n positions as evaluator rows (each row points at its own pod for the maze).
static ar_status stage_rows(ar_engine* e, const ar_game_pod* games, int n) {
  ar_status s = validate_games(e, games, n);
  if (s) return s;
  e->cur->resident_valid = false;  // the staging below reuses the upload buffers
  if (n > e->cur->cap_games) {
    cudaFree(e->cur->d_games); cudaFree(e->cur->d_seeds); cudaFree(e->cur->d_summaries); cudaFree(e->cur->d_positions);
    e->cur->d_games = nullptr; e->cur->d_seeds = nullptr; e->cur->d_summaries = nullptr; e->cur->d_positions = nullptr;
    CK(cudaMalloc(&e->cur->d_games, (size_t)n * sizeof(ar_game_pod)));
    CK(cudaMalloc(&e->cur->d_seeds, (size_t)n * sizeof(uint64_t)));
    CK(cudaMalloc(&e->cur->d_summaries, (size_t)n * sizeof(ar_game_summary)));
    e->cur->cap_games = n;
    e->cur->cap_stride = 0;
  }
  if (n > e->cap_rows) {
    cudaFree(e->d_rows); cudaFree(e->d_nn_out);
    e->d_rows = nullptr; e->d_nn_out = nullptr;
    CK(cudaMalloc(&e->d_rows, (size_t)n * sizeof(EvalRow)));
    CK(cudaMalloc(&e->d_nn_out, (size_t)n * 12 * sizeof(float)));
    e->cap_rows = n;
  }
  std::vector<EvalRow> rows(n);
  for (int i = 0; i < n; ++i) pod_to_row(games[i], (uint32_t)i, rows[i]);
  CK(cudaMemcpyAsync(e->cur->d_games, games, (size_t)n * sizeof(ar_game_pod), cudaMemcpyHostToDevice, e->cur->stream));
  CK(cudaMemcpyAsync(e->d_rows, rows.data(), (size_t)n * sizeof(EvalRow), cudaMemcpyHostToDevice, e->cur->stream));
  CK(cudaStreamSynchronize(e->cur->stream));
  return AR_OK;
}

static void batch_free(BatchBuf& b) {
  cudaFree(b.d_games); cudaFree(b.d_seeds); cudaFree(b.d_summaries); cudaFree(b.d_positions);
  cudaFree(b.d_next); cudaFree(b.d_counters); cudaFree(b.d_error);
  cudaFree(b.d_dense); cudaFree(b.d_offsets);
  if (b.h_dense) cudaFreeHost(b.h_dense);
  if (b.h_games_pinned) cudaFreeHost(b.h_games_pinned);
  if (b.h_seeds_pinned) cudaFreeHost(b.h_seeds_pinned);
  if (b.ev0) cudaEventDestroy(b.ev0);
  if (b.ev1) cudaEventDestroy(b.ev1);
  if (b.stream) cudaStreamDestroy(b.stream);
  b = BatchBuf();
}

static cudaError_t batch_init(BatchBuf& b) {
  cudaError_t ce = cudaStreamCreateWithFlags(&b.stream, cudaStreamNonBlocking);
  if (ce == cudaSuccess) ce = cudaEventCreate(&b.ev0);
  if (ce == cudaSuccess) ce = cudaEventCreate(&b.ev1);
  if (ce == cudaSuccess) ce = cudaMalloc(&b.d_next, sizeof(unsigned int));
  if (ce == cudaSuccess) ce = cudaMalloc(&b.d_counters, 8 * sizeof(unsigned long long));
  if (ce == cudaSuccess) ce = cudaMalloc(&b.d_error, sizeof(int));
  return ce;
}

// SelfPlayStats::from_games (selfplay.rs:212-224)
static void fill_game_stats(ar_stats* stats, const ar_game_summary* summaries, int n) {
  stats->total_games = 0; stats->total_positions = 0; stats->total_simulations = 0;
  stats->total_nn_evals = 0; stats->total_terminals = 0; stats->total_collisions = 0;
  stats->total_cheese_collected = 0; stats->total_cheese_available = 0;
  stats->p1_wins = stats->p2_wins = stats->draws = 0;
  stats->min_turns = 0xffffffffu;
  stats->max_turns = 0;
  for (int i = 0; i < n; ++i) {
    const ar_game_summary& g = summaries[i];
    stats->total_games += 1;
    stats->total_positions += g.n_positions;
    stats->total_simulations += g.total_simulations;
    stats->total_nn_evals += g.total_nn_evals;
    stats->total_terminals += g.total_terminals;
    stats->total_collisions += g.total_collisions;
    stats->total_cheese_collected += g.final_p1_score + g.final_p2_score;
    stats->total_cheese_available += g.cheese_available;
    stats->min_turns = std::min(stats->min_turns, g.n_positions);
    stats->max_turns = std::max(stats->max_turns, g.n_positions);
    if (g.result == 1) stats->p1_wins++; else if (g.result == 2) stats->p2_wins++; else stats->draws++;
  }
  if (n == 0) stats->min_turns = 0;
}

extern "C" {

uint32_t ar_abi_version(void) { return AR_ABI_VERSION; }

const char* ar_last_error(const ar_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

ar_status ar_engine_create(const ar_engine_cfg* cfg, ar_engine** out) {
  if (!cfg || !out) { g_create_error = "cfg/out is NULL"; return AR_ERR_INVALID_ARG; }
  if (cfg->abi_version != AR_ABI_VERSION) { g_create_error = "ABI version mismatch"; return AR_ERR_INVALID_ARG; }
  ar_engine* e = new ar_engine();
  e->cfg = *cfg;
  auto fail = [&](ar_status s, const std::string& m) {
    g_create_error = m;
    ar_engine_destroy(e);
    return s;
  };
  if (cfg->concurrent_games == 0) return fail(AR_ERR_INVALID_ARG, "concurrent_games must be > 0");
  if (cfg->tree_engine > AR_TREE_HALF) return fail(AR_ERR_INVALID_ARG, "tree_engine must be AR_TREE_WARP, AR_TREE_THREAD or AR_TREE_HALF");
  if (cfg->max_cells == 0 || cfg->max_cells > AR_MAX_CELLS) return fail(AR_ERR_UNSUPPORTED, "max_cells must be in [1, 256]");
  if (cfg->max_cells > 64 && cfg->tree_engine != AR_TREE_THREAD)
    return fail(AR_ERR_UNSUPPORTED, "boards over 64 cells need tree_engine = AR_TREE_THREAD (the warp engine and the evaluators keep a one-word cheese bitboard)");
  if (cfg->max_batch_size == 0 || cfg->max_batch_size > MAX_BATCH) return fail(AR_ERR_INVALID_ARG, "max_batch_size must be in [1, 64]");
  if (cfg->max_turns == 0 || cfg->max_turns > 250) return fail(AR_ERR_UNSUPPORTED, "max_turns must be in [1, 250] in this build (8-bit path depth)");
  cudaError_t ce = cudaSetDevice(cfg->device);
  if (ce != cudaSuccess) return fail(AR_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(ce));
  e->device = cfg->device;
  e->n_slots = cfg->concurrent_games;
  // Node capacity per tree: a tree can keep growing through tree reuse, by at most `simulations`
  // nodes per move, so max_turns * max_simulations + 2 nodes can never overflow.  The thread-per-tree
  // engine pages its pools (capacity = page-table length, memory is committed as trees grow); the
  // NN-guided engine sizes fixed pools when it is first used (ensure_nn_pools).
  {
    const uint64_t worst = (uint64_t)cfg->max_turns * std::max<uint32_t>(cfg->max_simulations, 1) + 2;
    uint64_t cap = cfg->pool_nodes ? cfg->pool_nodes : worst;
    cap = std::min<uint64_t>(cap, (1ull << PATH_NODE_BITS) - 1);
    if (cap < 64) return fail(AR_ERR_INVALID_ARG, "pool_nodes must be >= 64");
    e->node_cap = (uint32_t)cap;
    e->tt_pt_stride = (uint32_t)((cap + tt::PAGE_NODES - 1) / tt::PAGE_NODES);
    e->tt_slots = (cfg->concurrent_games + TB_BLOCK - 1) / TB_BLOCK * TB_BLOCK;
  }
  e->max_depth = cfg->max_turns + 1;
  e->path_stride = e->max_depth + 1;
  e->batch_cap = cfg->max_batch_size;
  size_t smem = 4 * warp_smem_bytes(e->max_depth, e->batch_cap);
  if (smem > 227 * 1024) return fail(AR_ERR_UNSUPPORTED, "max_turns/max_batch_size need more than 227 KB of shared memory");
#define CKC(call)                                                                      \
  do {                                                                                 \
    cudaError_t _e = (call);                                                           \
    if (_e != cudaSuccess) return fail(AR_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e)); \
  } while (0)
  CKC(batch_init(e->main));
  CKC(cudaStreamCreateWithFlags(&e->stream2, cudaStreamNonBlocking));
  CKC(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
  CKC(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
  e->coll_len = (uint32_t)std::min<uint64_t>((uint64_t)e->node_cap + 2, 1u << 20);
  CKC(cudaMalloc(&e->coll_table, (size_t)e->coll_len * sizeof(uint16_t)));
  CKC(cudaHostAlloc(&e->h_progress, sizeof(ar_progress), cudaHostAllocMapped));
  CKC(cudaHostGetDevicePointer(&e->d_progress, e->h_progress, 0));
  CKC(cudaFuncSetAttribute(selfplay_uniform_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CKC(cudaFuncSetAttribute(selfplay_uniform_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem / 4));
  const size_t hsmem = 2 * hw::half_smem_bytes(e->max_depth, e->batch_cap);  // one warp of the two-trees-per-warp kernel
  if (4 * hsmem <= 227 * 1024) {
    CKC(cudaFuncSetAttribute(selfplay_half_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * hsmem)));
    CKC(cudaFuncSetAttribute(selfplay_half_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsmem));
  } else if (cfg->tree_engine == AR_TREE_HALF) {
    return fail(AR_ERR_UNSUPPORTED, "max_turns/max_batch_size need too much shared memory for two trees per warp");
  }
  CKC(cudaFuncSetAttribute(nn_step_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CKC(cudaFuncSetAttribute(nn_step_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem / 4));
#undef CKC
  *out = e;
  return AR_OK;
}

void ar_stream_close(ar_engine* e);

void ar_engine_destroy(ar_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  ar_stream_close(e);
  cudaFree(e->pools); cudaFree(e->path_bufs); cudaFree(e->remaps); cudaFree(e->coll_table);
  cudaFree(e->tt_arena); cudaFree(e->tt_bitmap); cudaFree(e->tt_page_tables); cudaFree(e->tt_arrs);
  cudaFree(e->d_slot_bitmap);
  batch_free(e->main);
  cudaFree(e->d_search);
  cudaFree(e->d_rows); cudaFree(e->d_nn_out);
  cudaFree(e->d_maze_tab);
  cudaFree(e->d_slots); cudaFree(e->d_tp_store); cudaFree(e->d_queue); cudaFree(e->d_queue_out); cudaFree(e->d_n_rows);
  cudaFree(e->d_cache); cudaFree(e->d_key_store); cudaFree(e->d_board_store);
  delete e->eval;
  if (e->h_progress) cudaFreeHost(e->h_progress);
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  if (e->ev_join) cudaEventDestroy(e->ev_join);
  if (e->ev_base) cudaEventDestroy(e->ev_base);
  if (e->stream2) cudaStreamDestroy(e->stream2);
  delete e;
}

ar_status ar_engine_load_weights(ar_engine* e, int32_t arch, int32_t width, int32_t height,
                                 const ar_tensor_desc* tensors, int32_t n_tensors) {
  if (!e) return AR_ERR_INVALID_ARG;
  CK(cudaSetDevice(e->device));
  delete e->eval;
  e->eval = nullptr;
  e->arch = AR_ARCH_UNIFORM;
  if (arch == AR_ARCH_UNIFORM) return AR_OK;
  if (arch != AR_ARCH_MLP && arch != AR_ARCH_SYMMETRIC && arch != AR_ARCH_CNN) {
    e->err = "unknown architecture " + std::to_string(arch);
    return AR_ERR_UNSUPPORTED;
  }
  if (width <= 0 || height <= 0) { e->err = "bad board size"; return AR_ERR_INVALID_ARG; }
  if (width * height > 64 || e->cfg.max_cells > 64) {
    e->err = "the leaf evaluators support boards of up to 64 cells (engine max_cells must be <= 64 too)";
    return AR_ERR_UNSUPPORTED;
  }
  if (!tensors || n_tensors <= 0) { e->err = "no tensors"; return AR_ERR_INVALID_ARG; }
  LeafEvaluator* ev = arch == AR_ARCH_MLP ? static_cast<LeafEvaluator*>(new MlpModel())
                      : arch == AR_ARCH_SYMMETRIC ? make_symmetric_evaluator() : make_cnn_evaluator();
  int rc = ev->load(tensors, n_tensors, width, height, e->err);
  if (rc != AR_OK) { delete ev; return (ar_status)rc; }
  e->eval = ev;
  e->arch = arch;
  e->nn_width = width;
  e->nn_height = height;
  return AR_OK;
}

// CachedBackend::new(inner, capacity) (cached_backend.rs:62-70): `entries_per_tree` positions per resident
// tree, rounded up to a power of two and capped at 65536 and at 8 GiB in total; 0 disables the cache.
ar_status ar_engine_set_eval_cache(ar_engine* e, uint32_t entries_per_tree) {
  if (!e) return AR_ERR_INVALID_ARG;
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->cur->stream));
  cudaFree(e->d_cache);
  e->d_cache = nullptr;
  e->cache_entries = 0;
  if (entries_per_tree == 0) return AR_OK;
  uint32_t n = 1;
  while (n < entries_per_tree && n < 65536u) n <<= 1;
  while (n > 64 && (size_t)n * e->n_slots * sizeof(CacheEnt) > ((size_t)8 << 30)) n >>= 1;
  CK(cudaMalloc(&e->d_cache, (size_t)n * e->n_slots * sizeof(CacheEnt)));
  e->cache_entries = n;
  return AR_OK;
}


static void free_tt(ar_engine* e) {
  cudaFree(e->tt_arena); cudaFree(e->tt_bitmap); cudaFree(e->tt_page_tables); cudaFree(e->tt_arrs);
  e->tt_arena = nullptr; e->tt_bitmap = nullptr; e->tt_page_tables = nullptr; e->tt_arrs = nullptr;
  e->tt_n_pages = 0;
}
static void free_nn_pools(ar_engine* e) {
  cudaFree(e->pools); cudaFree(e->path_bufs); cudaFree(e->remaps);
  e->pools = nullptr; e->path_bufs = nullptr; e->remaps = nullptr;
  e->pool_nodes = 0;
}

// Paged arena of the thread-per-tree engine: every resident tree owns page `slot` for good, the rest
// is handed out on demand.  Sized for the worst case when that fits in 70 % of free HBM (it never
// does for production configurations: 32768 trees x 93 pages), otherwise everything that fits.
static ar_status ensure_tt(ar_engine* e) {
  if (e->tt_arena) return AR_OK;
  free_nn_pools(e);
  size_t free_b = 0, total_b = 0;
  CK(cudaMemGetInfo(&free_b, &total_b));
  const uint64_t worst = (uint64_t)e->tt_slots * e->tt_pt_stride + 8;
  uint64_t fit = (uint64_t)(0.70 * (double)free_b) / tt::PAGE_BYTES;
  if (const char* v = getenv("AR_TT_ARENA_GB"))  // profiling knob: a small arena keeps ncu's save / restore of device memory cheap
    fit = std::min<uint64_t>(fit, (uint64_t)(atof(v) * 1e9) / tt::PAGE_BYTES);
  const uint64_t pages = std::min(worst, fit);
  if (pages < (uint64_t)e->tt_slots + 4) {
    e->err = "not enough free device memory for " + std::to_string(e->tt_slots) + " resident trees";
    return AR_ERR_CUDA;
  }
  e->tt_n_pages = (uint32_t)pages;
  e->tt_bitmap_words = (e->tt_n_pages + 31) / 32;
  CK(cudaMalloc(&e->tt_arena, (size_t)pages * tt::PAGE_BYTES));
  CK(cudaMalloc(&e->tt_bitmap, (size_t)e->tt_bitmap_words * sizeof(uint32_t)));
  CK(cudaMalloc(&e->tt_page_tables, (size_t)e->tt_slots * e->tt_pt_stride * sizeof(uint32_t)));
  const bool big = e->cfg.max_cells > 64;  // boards over 64 cells: the four-word cheese bitboard instantiation
  CK(cudaMalloc(&e->tt_arrs, (size_t)e->tt_slots * (big ? sizeof(tt::TT<4>::TArr) : sizeof(tt::TT<1>::TArr))));
  CK(cudaFuncSetAttribute(selfplay_tb_kernel<3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tb_smem_bytes<1>()));
  CK(cudaFuncSetAttribute(selfplay_tb_kernel<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tb_smem_bytes<4>()));
  return AR_OK;
}

// Fixed pools of the warp-per-tree NN-guided engine.
static ar_status ensure_nn_pools(ar_engine* e) {
  if (e->pools) return AR_OK;
  free_tt(e);
  size_t free_b = 0, total_b = 0;
  CK(cudaMemGetInfo(&free_b, &total_b));
  const uint64_t fit = (uint64_t)(0.70 * (double)free_b) / e->n_slots / (sizeof(NodeRec) + sizeof(uint32_t));
  uint64_t pn = e->cfg.pool_nodes ? e->cfg.pool_nodes : std::max<uint64_t>(std::min<uint64_t>(e->node_cap, fit), 64);
  pn = std::min<uint64_t>(pn, (1ull << PATH_NODE_BITS) - 1);
  e->pool_nodes = (uint32_t)pn;
  CK(cudaMalloc(&e->pools, (size_t)e->n_slots * (size_t)pn * sizeof(NodeRec) + sizeof(NodeRec)));  // + slack: AR_TREE_HALF offsets odd pools
  CK(cudaMalloc(&e->path_bufs, (size_t)e->n_slots * e->batch_cap * e->path_stride * sizeof(uint32_t)));
  CK(cudaMalloc(&e->remaps, (size_t)e->n_slots * pn * sizeof(uint32_t)));
  return AR_OK;
}

static RunParams make_params(ar_engine* e, const ar_search_cfg* cfg) {
  RunParams p{};
  p.sp.c_puct = cfg->c_puct;
  p.sp.fpu_reduction = cfg->fpu_reduction;
  p.sp.force_k = cfg->force_k;
  p.sp.noise_epsilon = cfg->noise_epsilon;
  p.sp.noise_concentration = cfg->noise_concentration;
  p.sp.n_sims = cfg->simulations;
  p.sp.batch_size = cfg->batch_size;
  p.pools = e->pools;
  p.pool_nodes = e->pool_nodes;
  p.path_bufs = e->path_bufs;
  p.path_stride = e->path_stride;
  p.remaps = e->remaps;
  p.coll_table = e->coll_table;
  p.coll_table_len = e->coll_len;
  p.max_depth = e->max_depth;
  p.batch_cap = e->batch_cap;
  p.n_slots = e->n_slots;
  p.next_game = e->cur->d_next;
  p.counters = e->cur->d_counters;
  p.error_flag = e->cur->d_error;
  return p;
}

// (Re)build the evaluators' per-game maze table for the n games resident in e->cur->d_games.
static ar_status ensure_maze_table(ar_engine* e, int n) {
  if (n > e->cap_maze) {
    cudaFree(e->d_maze_tab);
    e->d_maze_tab = nullptr;
    CK(cudaMalloc(&e->d_maze_tab, (size_t)n * MAZE_TAB_STRIDE * sizeof(uint16_t)));
    e->cap_maze = n;
  }
  CK(build_maze_table(e->cur->d_games, n, e->d_maze_tab, e->cur->stream));
  return AR_OK;
}

static ar_status ensure_nn_buffers(ar_engine* e) {
  if (e->d_slots) return AR_OK;
  size_t max_rows = (size_t)e->n_slots * e->batch_cap;
  CK(cudaMalloc(&e->d_slots, (size_t)e->n_slots * sizeof(SlotState)));
  CK(cudaMalloc(&e->d_tp_store, max_rows * sizeof(TpEntry)));
  CK(cudaMalloc(&e->d_queue, max_rows * sizeof(EvalRow)));
  CK(cudaMalloc(&e->d_queue_out, max_rows * 12 * sizeof(float)));
  CK(cudaMalloc(&e->d_n_rows, 4 * sizeof(uint32_t)));  // [0] queue length of group 0, [1] done slots, [2] queue length of group 1
  CK(cudaMalloc(&e->d_key_store, max_rows * sizeof(GPack)));
  CK(cudaMalloc(&e->d_board_store, (size_t)e->n_slots * SM_TP));
  return AR_OK;
}

static ar_status launch_and_wait(ar_engine* e, RunParams& p, ar_progress* user_progress, float* ms) {
  for (const BatchBuf* b : e->sbufs)
    if (b->in_flight) {  // a blocking launch takes its tree slots by block index, streaming launches by bitmap
      e->err = "a streaming batch is in flight: collect it (ar_stream_collect / ar_stream_wait) before a blocking call";
      return AR_ERR_INVALID_ARG;
    }
  CK(cudaMemsetAsync(e->cur->d_next, 0, sizeof(unsigned int), e->cur->stream));
  CK(cudaMemsetAsync(e->cur->d_counters, 0, 8 * sizeof(unsigned long long), e->cur->stream));
  CK(cudaMemsetAsync(e->cur->d_error, 0, sizeof(int), e->cur->stream));
  memset((void*)e->h_progress, 0, sizeof(ar_progress));
  p.progress = user_progress ? e->d_progress : nullptr;
  size_t smem = 4 * warp_smem_bytes(e->max_depth, e->batch_cap);
  int slots = std::min<int>(e->n_slots, std::max(p.n_games, 1));
  p.n_slots = slots;
  const bool nn = e->arch != AR_ARCH_UNIFORM;
  const bool tt_run = !nn && e->cfg.tree_engine == AR_TREE_THREAD;
  if (tt_run) {
    ar_status s = ensure_tt(e);
    if (s) return s;
    // every launch starts with only the trees' own first pages taken: bits [0, tt_slots)
    CK(cudaMemsetAsync(e->tt_bitmap, 0, (size_t)e->tt_bitmap_words * sizeof(uint32_t), e->cur->stream));
    CK(cudaMemsetAsync(e->tt_bitmap, 0xff, (size_t)e->tt_slots / 8, e->cur->stream));
  }
  if (!tt_run) {
    ar_status s = ensure_nn_pools(e);
    if (s) return s;
    p.pools = e->pools; p.pool_nodes = e->pool_nodes; p.path_bufs = e->path_bufs; p.remaps = e->remaps;
  }
  if (nn) {
    ar_status s = ensure_nn_buffers(e);
    if (s) return s;
    s = ensure_maze_table(e, p.n_games);
    if (s) return s;
  }
  CK(cudaEventRecord(e->cur->ev0, e->cur->stream));
  if (!nn && !tt_run && e->cfg.tree_engine == AR_TREE_HALF) {
    selfplay_half_kernel<4><<<(slots + 7) / 8, 128, 8 * hw::half_smem_bytes(e->max_depth, e->batch_cap), e->cur->stream>>>(p);
    CK(cudaGetLastError());
    e->cur->launches += 1;
  } else if (!nn && !tt_run) {
    selfplay_uniform_kernel<4><<<(slots + 3) / 4, 128, smem, e->cur->stream>>>(p);
    CK(cudaGetLastError());
    e->cur->launches += 1;
  } else if (tt_run) {
    tt::Ctx c{};
    c.arena = e->tt_arena;
    c.page_bitmap = e->tt_bitmap;
    c.n_pages = e->tt_n_pages;
    c.bitmap_words = e->tt_bitmap_words;
    c.page_tables = e->tt_page_tables;
    c.pt_stride = e->tt_pt_stride;
    c.coll_table = e->coll_table;
    c.coll_len = e->coll_len;
    c.sp.c_puct = p.sp.c_puct; c.sp.fpu_reduction = p.sp.fpu_reduction; c.sp.force_k = p.sp.force_k;
    c.sp.noise_epsilon = p.sp.noise_epsilon; c.sp.noise_concentration = p.sp.noise_concentration;
    c.sp.n_sims = p.sp.n_sims; c.sp.batch_size = p.sp.batch_size;
    c.games = p.games; c.seeds = p.seeds; c.n_games = p.n_games;
    c.next_game = e->cur->d_next;
    c.summaries = p.summaries; c.positions = p.positions; c.pos_stride = p.pos_stride;
    c.search_out = p.search_out; c.search_only = p.search_only;
    { const char* v = getenv("AR_TT_MAX_MOVES"); c.max_moves = v ? atoi(v) : 0; }
    c.counters = e->cur->d_counters;
    c.error_flag = e->cur->d_error;
    c.progress = p.progress;
    const int tslots = (std::min<int>((int)e->tt_slots, std::max(p.n_games, 1)) + TB_BLOCK - 1) / TB_BLOCK * TB_BLOCK;
    // AR_TT_KERNEL: 0 = lane-bound (a tree stays on its lane; the faster of the two, default), 1 = block-sorted (tree
    // state in shared memory, trees sorted by phase every iteration); AR_TT_WARPS_PER_SM: register budget of variant 0
    static const int variant = [] { const char* v = getenv("AR_TT_KERNEL"); return v ? atoi(v) : 0; }();
    static const int minb = [] { const char* v = getenv("AR_TT_WARPS_PER_SM"); return v ? atoi(v) : 12; }();
    const bool big = e->cfg.max_cells > 64;
    const int nb = tslots / TT_BLOCK;
    if (big) {
      if (variant == 1) selfplay_tb_kernel<2, 4><<<tslots / TB_BLOCK, TB_BLOCK, tb_smem_bytes<4>(), e->cur->stream>>>(c, (tt::TT<4>::TArr*)e->tt_arrs);
      else selfplay_tt_kernel<8, 4><<<nb, TT_BLOCK, TT_BLOCK * 4 * tt::TT<4>::MAZE_WORDS, e->cur->stream>>>(c, tslots);
    } else if (variant == 1) {
      selfplay_tb_kernel<3, 1><<<tslots / TB_BLOCK, TB_BLOCK, tb_smem_bytes<1>(), e->cur->stream>>>(c, (tt::TT<1>::TArr*)e->tt_arrs);
    } else {
      const size_t sm = TT_BLOCK * 4 * tt::TT<1>::MAZE_WORDS;
      if (minb >= 20) selfplay_tt_kernel<20, 1><<<nb, TT_BLOCK, sm, e->cur->stream>>>(c, tslots);
      else if (minb >= 16) selfplay_tt_kernel<16, 1><<<nb, TT_BLOCK, sm, e->cur->stream>>>(c, tslots);
      else if (minb >= 12) selfplay_tt_kernel<12, 1><<<nb, TT_BLOCK, sm, e->cur->stream>>>(c, tslots);
      else selfplay_tt_kernel<8, 1><<<nb, TT_BLOCK, sm, e->cur->stream>>>(c, tslots);
    }
    CK(cudaGetLastError());
    e->cur->launches += 1;
  } else {
    NnParams q{};
    q.slots = e->d_slots;
    q.tp_store = e->d_tp_store;
    q.rows = e->d_queue;
    q.n_rows = e->d_n_rows;
    q.nn_out = e->d_queue_out;
    q.done_slots = e->d_n_rows + 1;
    q.max_rows = (uint32_t)((size_t)slots * e->batch_cap);
    q.key_store = e->d_key_store;
    q.board_store = e->d_board_store;
    // With the evaluation cache most batches complete without the evaluator; chaining them inside a step
    // only lengthens the step's critical path (measured: CNN config 9.1 s with 1 round, 9.7 s with 3, 16.6 s with 12).
    q.max_iters = e->cache_entries ? 1 : 3;
    if (e->cache_entries) {
      // a run starts with an empty cache: game indices and weights may have changed since the last one
      q.cache = e->d_cache;
      q.cache_mask = e->cache_entries - 1;
      CK(cudaMemsetAsync(e->d_cache, 0, (size_t)slots * e->cache_entries * sizeof(CacheEnt), e->cur->stream));
    }
    CK(cudaMemsetAsync(e->d_n_rows, 0, 4 * sizeof(uint32_t), e->cur->stream));
    nn_init_slots_kernel<<<(slots + 127) / 128, 128, 0, e->cur->stream>>>(e->d_slots, slots);
    CK(cudaGetLastError());
    e->cur->launches += 1;
    // With more slots than one wave of warps (148 SMs x 28), the slots are split into two groups that step on
    // two streams: a step is bulk-synchronous (tree kernel, then evaluator), so while one group's stragglers
    // finish or its leaves are being scored, the other group's tree kernel fills the SMs (config 3, 16384
    // slots: +7 %).  A single wave is better left whole (4096 slots: two groups of 2048 cost 9 %).  Each
    // group has its own evaluation queue; games, trees, caches and counters are shared arrays indexed by
    // the global slot.
    const int n_groups = slots >= 8192 ? 2 : 1;
    NnParams qs[2] = {q, q};
    cudaStream_t streams[2] = {e->cur->stream, e->stream2};
    for (int g = 0; g < n_groups; ++g) {
      const int s0 = n_groups == 2 ? ((slots / 2 + 3) & ~3) : slots;
      qs[g].slot_begin = g ? s0 : 0;
      qs[g].slot_end = g ? slots : s0;
      qs[g].rows = e->d_queue + (size_t)qs[g].slot_begin * e->batch_cap;
      qs[g].nn_out = e->d_queue_out + (size_t)qs[g].slot_begin * e->batch_cap * 12;
      qs[g].n_rows = e->d_n_rows + (g ? 2 : 0);
      qs[g].max_rows = (uint32_t)((size_t)(qs[g].slot_end - qs[g].slot_begin) * e->batch_cap);
    }
    // 32 steps (queue reset, tree step, evaluator; both groups) are captured once into a CUDA graph and
    // replayed: the step kernels are short (0.3-1 ms) and depend on each other, so launch gaps would otherwise
    // show.  The host reads one counter per replay to know when every game is finished.
    const int check_every = 32;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t gexec = nullptr;
    cudaError_t cap = cudaStreamBeginCapture(e->cur->stream, cudaStreamCaptureModeThreadLocal);
    if (cap == cudaSuccess) {
      cudaError_t in_cap = cudaSuccess;
      if (n_groups == 2) {
        in_cap = cudaEventRecord(e->ev_fork, e->cur->stream);
        if (in_cap == cudaSuccess) in_cap = cudaStreamWaitEvent(e->stream2, e->ev_fork, 0);
      }
      for (int it = 0; it < check_every && in_cap == cudaSuccess; ++it) {
        for (int g = 0; g < n_groups && in_cap == cudaSuccess; ++g) {
          const NnParams& qg = qs[g];
          in_cap = cudaMemsetAsync(qg.n_rows, 0, sizeof(uint32_t), streams[g]);
          if (in_cap != cudaSuccess) break;
          static const int nn_wpb = [] { const char* v = getenv("AR_NN_WPB"); return v ? atoi(v) : 1; }();
          if (nn_wpb == 4) nn_step_kernel<4><<<(qg.slot_end - qg.slot_begin + 3) / 4, 128, smem, streams[g]>>>(p, qg);
          else nn_step_kernel<1><<<qg.slot_end - qg.slot_begin, 32, smem / 4, streams[g]>>>(p, qg);
          in_cap = cudaGetLastError();
          if (in_cap != cudaSuccess) break;
          in_cap = e->eval->forward(qg.rows, qg.n_rows, (int)qg.max_rows, p.games, e->d_maze_tab,
                                    const_cast<float*>(qg.nn_out), e->cur->d_error, streams[g]);
        }
      }
      if (n_groups == 2 && in_cap == cudaSuccess) {
        in_cap = cudaEventRecord(e->ev_join, e->stream2);
        if (in_cap == cudaSuccess) in_cap = cudaStreamWaitEvent(e->cur->stream, e->ev_join, 0);
      }
      cap = cudaStreamEndCapture(e->cur->stream, &graph);
      if (in_cap != cudaSuccess) cap = in_cap;
      if (cap == cudaSuccess) cap = cudaGraphInstantiate(&gexec, graph, 0);
    }
    if (cap != cudaSuccess) {
      if (graph) cudaGraphDestroy(graph);
      e->err = std::string("CUDA graph capture of the NN step failed: ") + cudaGetErrorString(cap);
      return AR_ERR_CUDA;
    }
    ar_status loop_status = AR_OK;
    for (;;) {
      cudaError_t le = cudaGraphLaunch(gexec, e->cur->stream);
      e->cur->launches += 2 * check_every * n_groups;
      e->nn_steps += check_every;
      uint32_t h[2] = {0, 0};
      int herr = 0;
      if (le == cudaSuccess) le = cudaMemcpyAsync(h, e->d_n_rows, sizeof(h), cudaMemcpyDeviceToHost, e->cur->stream);
      if (le == cudaSuccess) le = cudaMemcpyAsync(&herr, e->cur->d_error, sizeof(int), cudaMemcpyDeviceToHost, e->cur->stream);
      if (le == cudaSuccess) le = cudaStreamSynchronize(e->cur->stream);
      if (le != cudaSuccess) {
        e->err = std::string("NN step: ") + cudaGetErrorString(le);
        loop_status = AR_ERR_CUDA;
        break;
      }
      if (user_progress) memcpy((void*)user_progress, (const void*)e->h_progress, sizeof(ar_progress));
      if (herr != 0 || (int)h[1] >= slots) break;
    }
    cudaGraphExecDestroy(gexec);
    cudaGraphDestroy(graph);
    if (loop_status != AR_OK) return loop_status;
  }
  CK(cudaEventRecord(e->cur->ev1, e->cur->stream));
  if (user_progress && !nn) {
    while (cudaEventQuery(e->cur->ev1) == cudaErrorNotReady) {
      memcpy((void*)user_progress, (const void*)e->h_progress, sizeof(ar_progress));
      std::this_thread::sleep_for(std::chrono::milliseconds(2));
    }
  }
  CK(cudaStreamSynchronize(e->cur->stream));
  if (user_progress) memcpy((void*)user_progress, (const void*)e->h_progress, sizeof(ar_progress));
  CK(cudaEventElapsedTime(ms, e->cur->ev0, e->cur->ev1));
  int herr = 0;
  CK(cudaMemcpy(&herr, e->cur->d_error, sizeof(int), cudaMemcpyDeviceToHost));
  if (herr != 0) {
    e->err = "device reported status " + std::to_string(herr) +
             (herr == AR_ERR_POOL_OVERFLOW ? " (node pool / depth stack / eval queue exhausted: raise pool_nodes or max_turns)"
              : herr == AR_ERR_NONFINITE ? " (the evaluator produced a non-finite value)" : "");
    return (ar_status)herr;
  }
  return AR_OK;
}

ar_status ar_search_batch(ar_engine* e, const ar_game_pod* games, int32_t n, const ar_search_cfg* cfg,
                          const uint64_t* seeds, ar_search_result* out) {
  if (!e) return AR_ERR_INVALID_ARG;
  CK(cudaSetDevice(e->device));
  ar_status s = validate_cfg(e, cfg);
  if (s) return s;
  s = validate_games(e, games, n);
  if (s) return s;
  if (n == 0) return AR_OK;
  if (!seeds || !out) { e->err = "seeds/out is NULL"; return AR_ERR_INVALID_ARG; }
  if ((uint64_t)cfg->simulations + 2 > e->node_cap) { e->err = "simulations exceed pool_nodes"; return AR_ERR_POOL_OVERFLOW; }
  if (e->arch != AR_ARCH_UNIFORM)
    for (int i = 0; i < n; ++i)
      if (games[i].width != e->nn_width || games[i].height != e->nn_height) {
        e->err = "game board size does not match the loaded evaluator";
        return AR_ERR_INVALID_ARG;
      }
  s = ensure_coll_table(e, *cfg);
  if (s) return s;
  e->cur->resident_valid = false;  // the search reuses the upload buffers
  if (n > e->cur->cap_games) {
    cudaFree(e->cur->d_games); cudaFree(e->cur->d_seeds); cudaFree(e->cur->d_summaries);
    e->cur->d_games = nullptr; e->cur->d_seeds = nullptr; e->cur->d_summaries = nullptr;
    CK(cudaMalloc(&e->cur->d_games, (size_t)n * sizeof(ar_game_pod)));
    CK(cudaMalloc(&e->cur->d_seeds, (size_t)n * sizeof(uint64_t)));
    CK(cudaMalloc(&e->cur->d_summaries, (size_t)n * sizeof(ar_game_summary)));
    e->cur->cap_games = n;
    e->cur->cap_stride = 0;
  }
  if (n > e->cap_search) {
    cudaFree(e->d_search);
    e->d_search = nullptr;
    CK(cudaMalloc(&e->d_search, (size_t)n * sizeof(ar_search_result)));
    e->cap_search = n;
  }
  CK(cudaMemcpyAsync(e->cur->d_games, games, (size_t)n * sizeof(ar_game_pod), cudaMemcpyHostToDevice, e->cur->stream));
  CK(cudaMemcpyAsync(e->cur->d_seeds, seeds, (size_t)n * sizeof(uint64_t), cudaMemcpyHostToDevice, e->cur->stream));
  RunParams p = make_params(e, cfg);
  p.games = e->cur->d_games; p.seeds = e->cur->d_seeds; p.n_games = n;
  p.search_only = 1; p.search_out = e->d_search;
  float ms = 0;
  s = launch_and_wait(e, p, nullptr, &ms);
  if (s) return s;
  CK(cudaMemcpy(out, e->d_search, (size_t)n * sizeof(ar_search_result), cudaMemcpyDeviceToHost));
  return AR_OK;
}

ar_status ar_selfplay_upload(ar_engine* e, const ar_game_pod* games, int32_t n, const uint64_t* seeds) {
  if (!e) return AR_ERR_INVALID_ARG;
  CK(cudaSetDevice(e->device));
  ar_status s = validate_games(e, games, n);
  if (s) return s;
  if (n > 0 && !seeds) { e->err = "seeds is NULL"; return AR_ERR_INVALID_ARG; }
  int stride = 1;
  for (int i = 0; i < n; ++i) stride = std::max<int>(stride, games[i].max_turns);
  if (n > e->cur->cap_games || stride > e->cur->cap_stride) {
    cudaFree(e->cur->d_games); cudaFree(e->cur->d_seeds); cudaFree(e->cur->d_summaries); cudaFree(e->cur->d_positions);
    e->cur->d_games = nullptr; e->cur->d_seeds = nullptr; e->cur->d_summaries = nullptr; e->cur->d_positions = nullptr;
    int cap = std::max(n, 1);
    CK(cudaMalloc(&e->cur->d_games, (size_t)cap * sizeof(ar_game_pod)));
    CK(cudaMalloc(&e->cur->d_seeds, (size_t)cap * sizeof(uint64_t)));
    CK(cudaMalloc(&e->cur->d_summaries, (size_t)cap * sizeof(ar_game_summary)));
    CK(cudaMalloc(&e->cur->d_positions, (size_t)cap * stride * sizeof(ar_position_record)));
    e->cur->cap_games = cap;
    e->cur->cap_stride = stride;
  }
  if (n > 0) {
    CK(cudaMemcpyAsync(e->cur->d_games, games, (size_t)n * sizeof(ar_game_pod), cudaMemcpyHostToDevice, e->cur->stream));
    CK(cudaMemcpyAsync(e->cur->d_seeds, seeds, (size_t)n * sizeof(uint64_t), cudaMemcpyHostToDevice, e->cur->stream));
    CK(cudaStreamSynchronize(e->cur->stream));
  }
  e->cur->h_games.assign(games, games + n);
  e->cur->resident_valid = true;
  e->cur->n_resident = n;
  e->cur->resident_stride = e->cur->cap_stride;
  e->cur->h2d += (uint64_t)n * (sizeof(ar_game_pod) + sizeof(uint64_t));
  return AR_OK;
}

static ar_status run_resident(ar_engine* e, const ar_search_cfg* cfg, ar_progress* progress, ar_stats* stats) {
  if (!e->cur->resident_valid) {
    e->err = "no uploaded batch: ar_selfplay_upload must be the last call that staged games on this engine";
    return AR_ERR_INVALID_ARG;
  }
  ar_status s = validate_cfg(e, cfg);
  if (s) return s;
  s = ensure_coll_table(e, *cfg);
  if (s) return s;
  if (e->arch != AR_ARCH_UNIFORM)
    for (const ar_game_pod& g : e->cur->h_games)
      if (g.width != e->nn_width || g.height != e->nn_height) {
        e->err = "game board size does not match the loaded evaluator";
        return AR_ERR_INVALID_ARG;
      }
  int n = e->cur->n_resident;
  float ms = 0;
  e->cur->launches = 0;
  RunParams p = make_params(e, cfg);
  if (n > 0) {
    p.games = e->cur->d_games; p.seeds = e->cur->d_seeds; p.n_games = n;
    p.summaries = e->cur->d_summaries; p.positions = e->cur->d_positions; p.pos_stride = e->cur->resident_stride;
    p.search_only = 0;
    auto t0 = std::chrono::steady_clock::now();
    s = launch_and_wait(e, p, progress, &ms);
    if (s) return s;
    if (stats) stats->elapsed_secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  if (stats) {
    unsigned long long c[8] = {0};
    if (n > 0) CK(cudaMemcpy(c, e->cur->d_counters, sizeof(c), cudaMemcpyDeviceToHost));
    stats->device_ms = ms;
    stats->path_nodes = c[0];
    stats->new_nodes = c[1];
    stats->kernel_launches = e->cur->launches;
    stats->cache_hits = c[6];
    stats->cache_misses = c[7];
#ifdef AR_PHASE_TIMING
    fprintf(stderr, "[phase cycles] gather=%llu backup=%llu advance=%llu total=%llu\n", c[2], c[3], c[4], c[5]);
#endif
  }
  return AR_OK;
}

ar_status ar_selfplay_run_resident(ar_engine* e, const ar_search_cfg* cfg, ar_stats* stats) {
  if (!e) return AR_ERR_INVALID_ARG;
  CK(cudaSetDevice(e->device));
  if (stats) memset(stats, 0, sizeof(*stats));
  return run_resident(e, cfg, nullptr, stats);
}

// Pack the resident batch's records on the device: game-major, n_positions records per game.  Needs the
// game lengths on the host (`summaries` receives all summaries); `off` gets the record offset of every game.
static ar_status pack_records(ar_engine* e, ar_game_summary* summaries, int positions_stride,
                              std::vector<uint32_t>& off) {
  BatchBuf* b = e->cur;
  const int n = b->n_resident;
  CK(cudaMemcpy(summaries, b->d_summaries, (size_t)n * sizeof(ar_game_summary), cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; ++i)
    if (positions_stride > 0 && (int)summaries[i].n_positions > positions_stride) {
      e->err = "positions_stride smaller than a game's length";
      return AR_ERR_INVALID_ARG;
    }
  off.assign((size_t)n + 1, 0);
  for (int i = 0; i < n; ++i) off[i + 1] = off[i] + summaries[i].n_positions;
  const size_t total = off[n];
  if (n + 1 > b->cap_offsets) {
    cudaFree(b->d_offsets);
    b->d_offsets = nullptr;
    CK(cudaMalloc(&b->d_offsets, ((size_t)n + 1) * sizeof(uint32_t)));
    b->cap_offsets = n + 1;
  }
  if (total > b->cap_dense) {
    cudaFree(b->d_dense);
    if (b->h_dense) cudaFreeHost(b->h_dense);
    b->d_dense = nullptr; b->h_dense = nullptr;
    CK(cudaMalloc(&b->d_dense, total * sizeof(ar_position_record)));
    CK(cudaHostAlloc(&b->h_dense, total * sizeof(ar_position_record), cudaHostAllocDefault));
    b->cap_dense = total;
  }
  if (total > 0) {
    CK(cudaMemcpyAsync(b->d_offsets, off.data(), off.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, b->stream));
    compact_positions_kernel<<<(n + 7) / 8, 256, 0, b->stream>>>(b->d_positions, b->resident_stride, b->d_offsets, n, b->d_dense);
    CK(cudaGetLastError());
  }
  b->launches += 1;
  return AR_OK;
}

ar_status ar_selfplay_download(ar_engine* e, ar_game_summary* summaries, ar_position_record* positions,
                               int32_t positions_stride) {
  if (!e) return AR_ERR_INVALID_ARG;
  CK(cudaSetDevice(e->device));
  if (!e->cur->resident_valid) {
    e->err = "no uploaded batch to download (a search or evaluator call reused the buffers)";
    return AR_ERR_INVALID_ARG;
  }
  int n = e->cur->n_resident;
  if (n == 0) return AR_OK;
  if (!summaries || !positions) { e->err = "summaries/positions is NULL"; return AR_ERR_INVALID_ARG; }
  if (positions_stride < 1) { e->err = "positions_stride < 1"; return AR_ERR_INVALID_ARG; }
  // pack on device, one bulk copy into pinned staging, scatter into the caller's strided array
  std::vector<uint32_t> off;
  ar_status ps = pack_records(e, summaries, positions_stride, off);
  if (ps) return ps;
  const size_t total = off[n];
  if (total > 0) {
    CK(cudaMemcpyAsync(e->cur->h_dense, e->cur->d_dense, total * sizeof(ar_position_record), cudaMemcpyDeviceToHost, e->cur->stream));
    CK(cudaStreamSynchronize(e->cur->stream));
  }
  e->cur->d2h += (uint64_t)n * sizeof(ar_game_summary) + total * sizeof(ar_position_record);
  // scatter into the caller's strided array, split over a few host threads
  auto finish = [&](int lo, int hi) {
    for (int i = lo; i < hi; ++i) {
      ar_position_record* dst = positions + (size_t)i * positions_stride;
      if (summaries[i].n_positions)
        memcpy(dst, e->cur->h_dense + off[i], (size_t)summaries[i].n_positions * sizeof(ar_position_record));
    }
  };
  const int n_thr = n >= 4096 ? (int)std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency())) : 1;
  if (n_thr <= 1) {
    finish(0, n);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < n_thr; ++t) pool.emplace_back(finish, (int)((int64_t)n * t / n_thr), (int)((int64_t)n * (t + 1) / n_thr));
    for (auto& th : pool) th.join();
  }
  return AR_OK;
}

// The resident batch's records, packed, as DEVICE pointers (valid until the next call on this engine):
// what a multi-GPU run hands to ncclGather / all_gather instead of bouncing records through the host.
ar_status ar_selfplay_pack_device(ar_engine* e, ar_game_summary* summaries, const void** d_summaries,
                                  const void** d_records, uint64_t* n_records) {
  if (!e) return AR_ERR_INVALID_ARG;
  CK(cudaSetDevice(e->device));
  if (!e->cur->resident_valid) { e->err = "no uploaded batch to pack"; return AR_ERR_INVALID_ARG; }
  if (!summaries || !d_summaries || !d_records || !n_records) { e->err = "output is NULL"; return AR_ERR_INVALID_ARG; }
  *d_summaries = e->cur->d_summaries;
  *d_records = nullptr;
  *n_records = 0;
  if (e->cur->n_resident == 0) return AR_OK;
  std::vector<uint32_t> off;
  ar_status ps = pack_records(e, summaries, 0, off);
  if (ps) return ps;
  CK(cudaStreamSynchronize(e->cur->stream));
  *d_records = e->cur->d_dense;
  *n_records = off[e->cur->n_resident];
  return AR_OK;
}

ar_status ar_selfplay_run(ar_engine* e, const ar_game_pod* games, int32_t n, const ar_search_cfg* cfg,
                          const uint64_t* seeds, ar_game_summary* summaries, ar_position_record* positions,
                          int32_t positions_stride, ar_progress* progress, ar_stats* stats) {
  if (!e) return AR_ERR_INVALID_ARG;
  CK(cudaSetDevice(e->device));
  auto t0 = std::chrono::steady_clock::now();
  if (stats) memset(stats, 0, sizeof(*stats));
  if (progress) memset((void*)progress, 0, sizeof(*progress));
  e->cur->h2d = e->cur->d2h = 0;
  ar_status s = validate_cfg(e, cfg);
  if (s) return s;
  s = ar_selfplay_upload(e, games, n, seeds);
  if (s) return s;
  s = run_resident(e, cfg, progress, stats);
  if (s) return s;
  s = ar_selfplay_download(e, summaries, positions, positions_stride);
  if (s) return s;
  if (stats) {
    fill_game_stats(stats, summaries, n);
    stats->elapsed_secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    stats->h2d_bytes = e->cur->h2d;
    stats->d2h_bytes = e->cur->d2h;
  }
  return AR_OK;
}

ar_status ar_encode_observations(ar_engine* e, const ar_game_pod* games, int32_t n, float* obs) {
  if (!e) return AR_ERR_INVALID_ARG;
  CK(cudaSetDevice(e->device));
  if (n == 0) return AR_OK;
  if (!obs) { e->err = "obs is NULL"; return AR_ERR_INVALID_ARG; }
  ar_status s = stage_rows(e, games, n);
  if (s) return s;
  int dim = 7 * games[0].width * games[0].height + 6;
  for (int i = 1; i < n; ++i)
    if (games[i].width != games[0].width || games[i].height != games[0].height) {
      e->err = "all games of one call must share the board size";
      return AR_ERR_INVALID_ARG;
    }
  float* d_obs = nullptr;
  CK(cudaMalloc(&d_obs, (size_t)n * dim * sizeof(float)));
  cudaError_t ce = encode_f32(e->d_rows, n, e->cur->d_games, dim, d_obs, e->cur->stream);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(obs, d_obs, (size_t)n * dim * sizeof(float), cudaMemcpyDeviceToHost, e->cur->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->cur->stream);
  cudaFree(d_obs);
  if (ce != cudaSuccess) { e->err = std::string("encode: ") + cudaGetErrorString(ce); return AR_ERR_CUDA; }
  return AR_OK;
}

ar_status ar_nn_forward(ar_engine* e, const ar_game_pod* games, int32_t n, float* policy_p1, float* policy_p2,
                        float* value_p1, float* value_p2) {
  if (!e) return AR_ERR_INVALID_ARG;
  CK(cudaSetDevice(e->device));
  if (e->arch == AR_ARCH_UNIFORM || !e->eval) { e->err = "no evaluator loaded (ar_engine_load_weights)"; return AR_ERR_NO_WEIGHTS; }
  if (n == 0) return AR_OK;
  if (!policy_p1 || !policy_p2 || !value_p1 || !value_p2) { e->err = "output is NULL"; return AR_ERR_INVALID_ARG; }
  ar_status s = stage_rows(e, games, n);
  if (s) return s;
  for (int i = 0; i < n; ++i)
    if (games[i].width != e->nn_width || games[i].height != e->nn_height) {
      e->err = "game " + std::to_string(i) + " does not match the evaluator's board size";
      return AR_ERR_INVALID_ARG;
    }
  CK(cudaMemsetAsync(e->cur->d_error, 0, sizeof(int), e->cur->stream));
  s = ensure_maze_table(e, n);
  if (s) return s;
  CK(e->eval->forward(e->d_rows, nullptr, n, e->cur->d_games, e->d_maze_tab, e->d_nn_out, e->cur->d_error, e->cur->stream));
  std::vector<float> out((size_t)n * 12);
  CK(cudaMemcpyAsync(out.data(), e->d_nn_out, out.size() * sizeof(float), cudaMemcpyDeviceToHost, e->cur->stream));
  int herr = 0;
  CK(cudaMemcpyAsync(&herr, e->cur->d_error, sizeof(int), cudaMemcpyDeviceToHost, e->cur->stream));
  CK(cudaStreamSynchronize(e->cur->stream));
  if (herr) { e->err = "evaluator produced a non-finite value"; return (ar_status)herr; }
  for (int i = 0; i < n; ++i) {
    memcpy(policy_p1 + i * 5, &out[(size_t)i * 12], 20);
    memcpy(policy_p2 + i * 5, &out[(size_t)i * 12 + 5], 20);
    value_p1[i] = out[(size_t)i * 12 + 10];
    value_p2[i] = out[(size_t)i * 12 + 11];
  }
  return AR_OK;
}

// -----------------------------------------------------------------------------------------
// Streaming self-play (continuous game feed): several batches in flight on one engine.
// Every batch is one persistent launch of selfplay_uniform_kernel on its own stream; launches share
// the engine's tree slots through the device-side group bitmap, so the next batch's blocks start as
// the previous batch's blocks finish their last (longest) games — throughput no longer depends on the
// number of games per call (a lone batch ends with a tail of long games running alone).
// Uniform-prior warp engine only.
// -----------------------------------------------------------------------------------------
static ar_status stream_buf(ar_engine* e, int32_t buffer, BatchBuf** out) {
  if (!e) return AR_ERR_INVALID_ARG;
  if (buffer < 0 || buffer >= (int)e->sbufs.size()) { e->err = "no such stream buffer (ar_stream_open)"; return AR_ERR_INVALID_ARG; }
  *out = e->sbufs[buffer];
  return AR_OK;
}

ar_status ar_stream_open(ar_engine* e, int32_t n_buffers, int32_t max_games, int32_t positions_stride) {
  if (!e) return AR_ERR_INVALID_ARG;
  CK(cudaSetDevice(e->device));
  if (n_buffers < 1 || n_buffers > 8 || max_games < 1 || positions_stride < 1) { e->err = "bad stream shape"; return AR_ERR_INVALID_ARG; }
  if (e->arch != AR_ARCH_UNIFORM || e->cfg.tree_engine == AR_TREE_THREAD) {
    e->err = "streaming is implemented for the uniform-prior warp engines (AR_TREE_WARP, AR_TREE_HALF)";
    return AR_ERR_UNSUPPORTED;
  }
  ar_stream_close(e);
  for (int i = 0; i < n_buffers; ++i) {
    BatchBuf* b = new BatchBuf();
    e->sbufs.push_back(b);
    CK(batch_init(*b));
    CK(cudaMalloc(&b->d_games, (size_t)max_games * sizeof(ar_game_pod)));
    CK(cudaMalloc(&b->d_seeds, (size_t)max_games * sizeof(uint64_t)));
    CK(cudaMalloc(&b->d_summaries, (size_t)max_games * sizeof(ar_game_summary)));
    CK(cudaMalloc(&b->d_positions, (size_t)max_games * positions_stride * sizeof(ar_position_record)));
    CK(cudaHostAlloc(&b->h_games_pinned, (size_t)max_games * sizeof(ar_game_pod), cudaHostAllocDefault));
    CK(cudaHostAlloc(&b->h_seeds_pinned, (size_t)max_games * sizeof(uint64_t), cudaHostAllocDefault));
    b->cap_games = max_games;
    b->cap_stride = positions_stride;
  }
  if (!e->ev_base) CK(cudaEventCreate(&e->ev_base));
  CK(cudaEventRecord(e->ev_base, e->main.stream));
  CK(cudaStreamSynchronize(e->main.stream));
  if (!e->d_slot_bitmap) {
    CK(cudaMalloc(&e->d_slot_bitmap, (size_t)((e->n_slots + 31) / 32 + 1) * sizeof(uint32_t)));
    CK(cudaMemset(e->d_slot_bitmap, 0, (size_t)((e->n_slots + 31) / 32 + 1) * sizeof(uint32_t)));
  }
  return AR_OK;
}

void ar_stream_close(ar_engine* e) {
  if (!e) return;
  for (BatchBuf* b : e->sbufs) {
    if (b->stream) cudaStreamSynchronize(b->stream);
    batch_free(*b);
    delete b;
  }
  e->sbufs.clear();
  e->cur = &e->main;
}

// Launch the resident batch of `buffer` (asynchronous).
ar_status ar_stream_launch(ar_engine* e, int32_t buffer, const ar_search_cfg* cfg) {
  BatchBuf* b = nullptr;
  ar_status s = stream_buf(e, buffer, &b);
  if (s) return s;
  CK(cudaSetDevice(e->device));
  if (b->in_flight) { e->err = "stream buffer is still in flight (ar_stream_wait / ar_stream_collect)"; return AR_ERR_INVALID_ARG; }
  if (!b->resident_valid) { e->err = "stream buffer holds no batch"; return AR_ERR_INVALID_ARG; }
  s = validate_cfg(e, cfg);
  if (s) return s;
  s = ensure_coll_table(e, *cfg);
  if (s) return s;
  s = ensure_nn_pools(e);
  if (s) return s;
  e->cur = b;
  RunParams p = make_params(e, cfg);
  e->cur = &e->main;
  p.pools = e->pools; p.pool_nodes = e->pool_nodes; p.path_bufs = e->path_bufs; p.remaps = e->remaps;
  p.games = b->d_games; p.seeds = b->d_seeds; p.n_games = b->n_resident;
  p.summaries = b->d_summaries; p.positions = b->d_positions; p.pos_stride = b->resident_stride;
  p.search_only = 0;
  const bool half_engine = e->cfg.tree_engine == AR_TREE_HALF;
  p.slot_bitmap = e->d_slot_bitmap;
  p.n_groups = half_engine ? (int)e->n_slots / 2 : (int)e->n_slots;  // one warp per block: one tree slot, or two
  p.n_slots = half_engine ? p.n_groups * 2 : p.n_groups;
  if (p.n_groups < 1) { e->err = "streaming needs at least one warp's worth of resident trees"; return AR_ERR_INVALID_ARG; }
  p.progress = nullptr;
  b->launches = 0;
  b->t_submit = std::chrono::steady_clock::now();
  CK(cudaMemsetAsync(b->d_next, 0, sizeof(unsigned int), b->stream));
  CK(cudaMemsetAsync(b->d_counters, 0, 8 * sizeof(unsigned long long), b->stream));
  CK(cudaMemsetAsync(b->d_error, 0, sizeof(int), b->stream));
  CK(cudaEventRecord(b->ev0, b->stream));
  if (b->n_resident > 0) {
    const size_t smem = warp_smem_bytes(e->max_depth, e->batch_cap);
    if (half_engine) {
      const int blocks = std::min(p.n_groups, (b->n_resident + 1) / 2);
      selfplay_half_kernel<1><<<blocks, 32, 2 * hw::half_smem_bytes(e->max_depth, e->batch_cap), b->stream>>>(p);
    } else {
      const int blocks = std::min(p.n_groups, b->n_resident);
      selfplay_uniform_kernel<1><<<blocks, 32, smem, b->stream>>>(p);
    }
    CK(cudaGetLastError());
    b->launches = 1;
  }
  CK(cudaEventRecord(b->ev1, b->stream));
  b->in_flight = true;
  return AR_OK;
}

// Copy a batch into `buffer` from host memory (asynchronous, through pinned staging) and launch it.
ar_status ar_stream_submit(ar_engine* e, int32_t buffer, const ar_game_pod* games, int32_t n,
                           const ar_search_cfg* cfg, const uint64_t* seeds) {
  BatchBuf* b = nullptr;
  ar_status s = stream_buf(e, buffer, &b);
  if (s) return s;
  CK(cudaSetDevice(e->device));
  if (b->in_flight) { e->err = "stream buffer is still in flight (ar_stream_collect)"; return AR_ERR_INVALID_ARG; }
  s = validate_games(e, games, n);
  if (s) return s;
  if (n > b->cap_games) { e->err = "batch larger than the stream's max_games"; return AR_ERR_INVALID_ARG; }
  if (n > 0 && !seeds) { e->err = "seeds is NULL"; return AR_ERR_INVALID_ARG; }
  for (int i = 0; i < n; ++i)
    if (games[i].max_turns > b->cap_stride) { e->err = "a game's max_turns exceeds the stream's positions_stride"; return AR_ERR_INVALID_ARG; }
  if (n > 0) {
    memcpy(b->h_games_pinned, games, (size_t)n * sizeof(ar_game_pod));
    memcpy(b->h_seeds_pinned, seeds, (size_t)n * sizeof(uint64_t));
    CK(cudaMemcpyAsync(b->d_games, b->h_games_pinned, (size_t)n * sizeof(ar_game_pod), cudaMemcpyHostToDevice, b->stream));
    CK(cudaMemcpyAsync(b->d_seeds, b->h_seeds_pinned, (size_t)n * sizeof(uint64_t), cudaMemcpyHostToDevice, b->stream));
  }
  b->n_resident = n;
  b->resident_stride = b->cap_stride;
  b->resident_valid = true;
  b->h2d = (uint64_t)n * (sizeof(ar_game_pod) + sizeof(uint64_t));
  b->d2h = 0;
  if (!cfg) return AR_OK;  // upload only: ar_stream_launch plays it
  return ar_stream_launch(e, buffer, cfg);
}

// Wait for the launch of `buffer`; stats (optional) carry device time and the roofline counters.
ar_status ar_stream_wait(ar_engine* e, int32_t buffer, ar_stats* stats) {
  BatchBuf* b = nullptr;
  ar_status s = stream_buf(e, buffer, &b);
  if (s) return s;
  CK(cudaSetDevice(e->device));
  if (!b->in_flight) { e->err = "stream buffer has no launch in flight"; return AR_ERR_INVALID_ARG; }
  CK(cudaStreamSynchronize(b->stream));
  b->in_flight = false;
  int herr = 0;
  CK(cudaMemcpy(&herr, b->d_error, sizeof(int), cudaMemcpyDeviceToHost));
  if (herr != 0) {
    e->err = "device reported status " + std::to_string(herr) + " (node pool / depth stack exhausted: raise pool_nodes or max_turns)";
    return (ar_status)herr;
  }
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, b->ev0, b->ev1));
    unsigned long long c[8] = {0};
    CK(cudaMemcpy(c, b->d_counters, sizeof(c), cudaMemcpyDeviceToHost));
    stats->device_ms = ms;
    stats->path_nodes = c[0];
    stats->new_nodes = c[1];
    stats->total_nn_evals = c[2];
    stats->total_terminals = c[3];
    stats->total_positions = c[4];
    stats->total_simulations = c[5];
    stats->total_games = (uint32_t)b->n_resident;
    stats->kernel_launches = b->launches;
    stats->elapsed_secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - b->t_submit).count();
  }
  return AR_OK;
}

// Wait for `buffer`, then download its records into host memory.
ar_status ar_stream_collect(ar_engine* e, int32_t buffer, ar_game_summary* summaries, ar_position_record* positions,
                            int32_t positions_stride, ar_stats* stats) {
  BatchBuf* b = nullptr;
  ar_status s = stream_buf(e, buffer, &b);
  if (s) return s;
  ar_stats local{};
  s = ar_stream_wait(e, buffer, &local);
  if (s) return s;
  e->cur = b;
  s = ar_selfplay_download(e, summaries, positions, positions_stride);
  e->cur = &e->main;
  if (s) return s;
  if (stats) {
    *stats = local;
    fill_game_stats(stats, summaries, b->n_resident);
    stats->kernel_launches = b->launches;
    stats->h2d_bytes = b->h2d;
    stats->d2h_bytes = b->d2h;
    stats->elapsed_secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - b->t_submit).count();
  }
  return AR_OK;
}

// Start and end of the last completed launch of `buffer` on the device clock, in ms since ar_stream_open
// (valid after ar_stream_wait / ar_stream_collect, until the buffer is launched again).
ar_status ar_stream_times(ar_engine* e, int32_t buffer, double* start_ms, double* end_ms) {
  BatchBuf* b = nullptr;
  ar_status s = stream_buf(e, buffer, &b);
  if (s) return s;
  if (!start_ms || !end_ms || b->in_flight) { e->err = "ar_stream_times: buffer in flight or NULL output"; return AR_ERR_INVALID_ARG; }
  float f0 = 0, f1 = 0;
  CK(cudaEventElapsedTime(&f0, e->ev_base, b->ev0));
  CK(cudaEventElapsedTime(&f1, e->ev_base, b->ev1));
  *start_ms = f0;
  *end_ms = f1;
  return AR_OK;
}

// Device time from the start of buffer `first`'s launch to the end of buffer `last`'s (both complete).
ar_status ar_stream_elapsed_ms(ar_engine* e, int32_t first, int32_t last, double* ms) {
  BatchBuf *a = nullptr, *b = nullptr;
  ar_status s = stream_buf(e, first, &a);
  if (s) return s;
  s = stream_buf(e, last, &b);
  if (s) return s;
  if (!ms) return AR_ERR_INVALID_ARG;
  float f = 0;
  CK(cudaEventElapsedTime(&f, a->ev0, b->ev1));
  *ms = f;
  return AR_OK;
}

}  // extern "C"
