// half_emul.cpp — TEST INFRASTRUCTURE ONLY.
//
// Compiles the device code of the two-trees-per-warp engine (alpharat_b200/csrc/mcts_half.cuh, and through it
// mcts_device.cuh) for the host behind tests/half_emul/simt_shim.h and runs one half of a warp — 16 lanes as 16
// cooperative fibers — through the per-half loop of selfplay_half_kernel (engine.cu), game after game.  It exists so
// that the lane mapping of the 256-byte record on a half, the width-16 / width-8 collectives, the FPU and prior
// tables, the compact shared-memory layout and the half-uniform control flow can be checked bit-for-bit against
// oracle/ in the CPU-only container (compute-sanitizer is closed on the GPU pool); the shim aborts when the lanes of
// the half do not execute the same sequence of collectives.  Even games run as lanes 0-15, odd games as lanes 16-31
// (hbase 16: the ballot shifts and the +128-byte pool of the second half).  The product never loads this file.
// Build: g++ -O2 -ffp-contract=off -shared (tests/half_emul_loader.py).
#include "simt_shim.h"

#include <vector>

#include "../../alpharat_b200/csrc/host_tables.hpp"
#include "../../alpharat_b200/csrc/mcts_half.cuh"

using namespace ar;

namespace {

struct Run {
  const ar_game_pod* games;
  const uint64_t* seeds;
  int n_games;
  SearchParams sp;
  int search_only;
  ar_game_summary* summaries;
  ar_position_record* positions;
  int pos_stride;
  ar_search_result* search_out;
  // per-half storage
  uint8_t* pool_bytes;  // pool_nodes records, 256-byte aligned, + 128 bytes of slack
  uint32_t pool_nodes;
  uint32_t* path_buf;
  uint32_t path_stride;
  uint32_t* remap;
  const uint16_t* coll_table;
  uint32_t coll_len;
  uint32_t max_depth, batch_cap;
  uint8_t* smem;
  const float* fpu_tab;
  int first_game, game_step_;  // this call plays games first_game, first_game + game_step_, ...
  // results
  unsigned long long path_nodes, new_nodes;
  uint32_t error;
};

// The loop of selfplay_half_kernel (engine.cu) for one half, with the game claim replaced by a fixed game list.
void half_body(int lane_in_half, void* arg) {
  Run& p = *static_cast<Run*>(arg);
  const int hbase = simt::g_half->hbase;
  const int half = hbase >> 4;
  int hl = lane_in_half;
  hw::HalfCtx cx;
  cx.fpu_tab = p.fpu_tab;
  cx.bind_half(p.smem, reinterpret_cast<NodeRec*>(p.pool_bytes + 128 * half), hl, p.max_depth, p.batch_cap);
  {
    const uint32_t lane_bits = ((uint32_t)(uintptr_t)cx.pool_lane >> 3) & 31u;  // as in the kernel
    hl = (int)(lane_bits & 15u);
    cx.hbase = (int)(lane_bits & 16u);
    if (hl != lane_in_half || cx.hbase != hbase) simt::fail("lane bits of the record pointer");
  }
  cx.path_buf = p.path_buf;
  cx.remap = p.remap;
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
  unsigned long long path_nodes = 0, new_nodes = 0;

  for (int gi = p.first_game; gi < p.n_games && cx.error == 0; gi += p.game_step_) {
    GState g;
    int turn = 0;
    hw::load_game(p.games + gi, cx, g, turn, hl);
    Rng rng = rng_seed(p.seeds[gi]);
    const int cheese_available = __popcll(g.cheese);
    hw::init_root(cx, g, hl);
    uint32_t n_pos = 0;
    unsigned long long tot_sims = 0, tot_nn = 0, tot_term = 0, tot_coll = 0;
    uint32_t remaining = sp.n_sims, nn = 0, term = 0, coll = 0;
    if (!p.search_only) {
      uint32_t* co = reinterpret_cast<uint32_t*>(p.summaries[gi].cheese_outcomes);
      for (int i = hl; i < AR_MAX_CELLS / 4; i += 16) co[i] = 0x02020202u;
      if (game_over(g, turn, cx.max_turns)) remaining = 0;
    }
    bool game_done = false;
    while (!game_done && cx.error == 0) {
      bool finished = !p.search_only && remaining == 0 && n_pos == 0 && game_over(g, turn, cx.max_turns);
      if (!finished) {
        if (remaining > 0) {
          const uint32_t bs = min(remaining, sp.batch_size);
          const uint32_t nn0 = nn, term0 = term;
          hw::simulate_batch_uniform(cx, sp, rng, g, turn, bs, nn, term, coll, p.coll_len, hl);
          uint32_t produced = (nn - nn0) + (term - term0);
          produced = produced > 1u ? produced : 1u;
          remaining = remaining > produced ? remaining - produced : 0u;
          if (cx.error) break;
        }
        if (remaining == 0) {
          float pol1[5], pol2[5];
          ar_position_record* pos = p.positions ? p.positions + (size_t)gi * p.pos_stride : nullptr;
          ar_search_result* rout = p.search_only ? (p.search_out + gi) : &pos[n_pos].search;
          ar_search_result res;
          hw::extract_result(cx, sp, hl, res);
          res.nn_evals = nn;
          res.terminals = term;
          res.collisions = coll;
          for (int a = 0; a < 5; ++a) { pol1[a] = res.policy_p1[a]; pol2[a] = res.policy_p2[a]; }
          const uint32_t tv = res.total_visits;
          if (hl == 0) *rout = res;
          if (p.search_only) {
            game_done = true;
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
              for (int t = 2; t < 8; ++t) cb[t] = 0;
            }
            n_pos += 1;
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
              finished = true;
            } else if (child != 0) {
              hw::compact_subtree(cx, child, hl);
            } else {
              hw::init_root(cx, g, hl);
            }
          }
        }
      }
      if (finished) {
        if (hl == 0) {
          ar_game_summary& s = p.summaries[gi];
          s.game_index = (uint32_t)gi;
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
        }
        game_done = true;
      }
    }
    path_nodes += cx.path_nodes;
    new_nodes += cx.new_nodes;
    cx.path_nodes = 0;
    cx.new_nodes = 0;
    hw::hsync();  // the next game's load_game rewrites the shared tables
  }
  if (hl == 0) {
    p.path_nodes += path_nodes;
    p.new_nodes += new_nodes;
    if (cx.error) p.error = cx.error;
  }
}

}  // namespace

extern "C" int half_emul_run(const ar_game_pod* games, int n, const ar_search_cfg* cfg, const uint64_t* seeds,
                             int pool_nodes, int search_only, ar_game_summary* summaries,
                             ar_position_record* positions, int stride, ar_search_result* search_out,
                             unsigned long long* counters /* [3]: path_nodes new_nodes collectives */) {
  if (n < 0 || pool_nodes < 2) return -1;
  Run p;
  memset(&p, 0, sizeof(p));
  p.games = games; p.seeds = seeds; p.n_games = n;
  p.sp.c_puct = cfg->c_puct; p.sp.fpu_reduction = cfg->fpu_reduction; p.sp.force_k = cfg->force_k;
  p.sp.noise_epsilon = cfg->noise_epsilon; p.sp.noise_concentration = cfg->noise_concentration;
  p.sp.n_sims = cfg->simulations; p.sp.batch_size = cfg->batch_size;
  p.search_only = search_only;
  p.summaries = summaries; p.positions = positions; p.pos_stride = stride; p.search_out = search_out;
  int max_turns = 1;
  for (int i = 0; i < n; ++i) max_turns = std::max<int>(max_turns, games[i].max_turns);
  p.max_depth = (uint32_t)max_turns + 1;   // engine.cu: max_depth = max_turns + 1
  p.path_stride = p.max_depth + 1;
  p.batch_cap = std::max<uint32_t>(cfg->batch_size, 1);
  p.pool_nodes = (uint32_t)pool_nodes;
  std::vector<uint8_t> pool((size_t)pool_nodes * sizeof(NodeRec) + 512);
  p.pool_bytes = pool.data() + ((256 - ((uintptr_t)pool.data() & 255)) & 255);
  std::vector<uint32_t> path_buf((size_t)p.batch_cap * p.path_stride), remap((size_t)pool_nodes);
  p.path_buf = path_buf.data();
  p.remap = remap.data();
  const uint32_t coll_len = (uint32_t)pool_nodes + 2;
  std::vector<uint16_t> coll = ar_host::collision_table(*cfg, coll_len);
  p.coll_table = coll.data();
  p.coll_len = coll_len;
  std::vector<uint8_t> smem(hw::half_smem_bytes(p.max_depth, p.batch_cap) + 64);
  p.smem = smem.data() + ((16 - ((uintptr_t)smem.data() & 15)) & 15);
  float tab[42];
  for (int i = 0; i < 36; ++i) tab[i] = hw::fpu_tab_entry(i / 6, i % 6);
  for (int i = 0; i < 6; ++i) tab[36 + i] = i ? 1.0f / (float)i : 0.0f;
  p.fpu_tab = tab;
  unsigned long long collectives = 0;
  for (int half = 0; half < 2; ++half) {  // even games as the first half of a warp, odd games as the second
    p.first_game = half;
    p.game_step_ = 2;
    collectives += simt::run_half(16 * half, half_body, &p);
  }
  if (counters) { counters[0] = p.path_nodes; counters[1] = p.new_nodes; counters[2] = collectives; }
  return (int)p.error;
}
