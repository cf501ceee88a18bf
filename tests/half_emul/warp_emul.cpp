// warp_emul.cpp — TEST INFRASTRUCTURE ONLY.
//
// The one-tree-per-warp device code (alpharat_b200/csrc/mcts_device.cuh: the uniform-prior path of
// selfplay_uniform_kernel, engine.cu) compiled for the host behind tests/half_emul/simt_shim.h and run as 32
// cooperative fibres, game after game.  Same purpose as half_emul.cpp: bit-exact parity against oracle/ in the CPU-only
// container, and a check that the 32 lanes of a warp execute the same sequence of collectives.  The product never
// loads this file.  Build: g++ -O2 -ffp-contract=off -shared (tests/half_emul_loader.py).
#include "simt_shim.h"

#include <vector>

#include "../../alpharat_b200/csrc/host_tables.hpp"
#include "../../alpharat_b200/csrc/mcts_device.cuh"

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
  uint8_t* pool_bytes;
  uint32_t pool_nodes;
  uint32_t* path_buf;
  uint32_t path_stride;
  uint32_t* remap;
  const uint16_t* coll_table;
  uint32_t coll_len;
  uint32_t max_depth, batch_cap;
  uint8_t* smem;
  unsigned long long path_nodes, new_nodes;
  uint32_t error;
};

// The loop of selfplay_uniform_kernel (engine.cu) for one warp, with the game claim replaced by a game counter.
void warp_body(int lane, void* arg) {
  Run& p = *static_cast<Run*>(arg);
  WarpCtx cx;
  cx.bind(p.smem, reinterpret_cast<NodeRec*>(p.pool_bytes), lane, p.max_depth, p.batch_cap);
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

  for (int gi = 0; gi < p.n_games; ++gi) {
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
      uint32_t remaining = sp.n_sims, nn = 0, term = 0, coll = 0;
      while (remaining > 0 && cx.error == 0) {
        uint32_t bs = min(remaining, sp.batch_size);
        uint32_t nn0 = nn, term0 = term;
        simulate_batch_uniform(cx, sp, rng, g, turn, bs, nn, term, coll, p.coll_len, lane);
        uint32_t produced = (nn - nn0) + (term - term0);
        produced = produced > 1u ? produced : 1u;
        remaining = remaining > produced ? remaining - produced : 0u;
      }
      if (cx.error) break;
      float pol1[5], pol2[5];
      ar_search_result* rout = p.search_only ? (p.search_out + gi) : &pos[n_pos].search;
      extract_and_store(cx, sp, lane, rout, nn, term, coll, pol1, pol2);
      if (p.search_only) break;
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
        uint32_t* cb = reinterpret_cast<uint32_t*>(pr.cheese);
        cb[0] = (uint32_t)g.cheese; cb[1] = (uint32_t)(g.cheese >> 32);
        for (int t = 2; t < 8; ++t) cb[t] = 0;
      }
      n_pos += 1;
      uint32_t rmeta = cx.pool[0].s[LANE_LINKS].y;
      int i = action_to_idx(meta_m1(rmeta), a1), j = action_to_idx(meta_m2(rmeta), a2);
      uint32_t child = reinterpret_cast<const uint32_t*>(&cx.pool[0].s[LANE_CHILD])[i * 5 + j];
      const uint64_t cheese_before = g.cheese;
      game_step(g, i, j, cx.steptbl());
      turn += 1;
      if (lane == 0) credit_cheese(p.summaries[gi], cheese_before, g);
      __syncwarp();
      if (game_over(g, turn, cx.max_turns)) break;
      if (child != 0) {
        compact_subtree(cx, child, lane);
      } else {
        cx.epoch += 1;
        init_root(cx, g, lane);
      }
    }
    if (cx.error) break;
    if (!p.search_only && lane == 0) {
      ar_game_summary& s = p.summaries[gi];
      s.game_index = (uint32_t)gi;
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
    }
    __syncwarp();  // the next game's load_game rewrites the shared tables
  }
  if (lane == 0) {
    p.path_nodes += cx.path_nodes;
    p.new_nodes += cx.new_nodes;
    if (cx.error) p.error = cx.error;
  }
}

}  // namespace

extern "C" int warp_emul_run(const ar_game_pod* games, int n, const ar_search_cfg* cfg, const uint64_t* seeds,
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
  p.max_depth = (uint32_t)max_turns + 1;
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
  std::vector<uint8_t> smem(warp_smem_bytes(p.max_depth, p.batch_cap) + 64);
  p.smem = smem.data() + ((16 - ((uintptr_t)smem.data() & 15)) & 15);
  const unsigned long long collectives = simt::run_half(0, warp_body, &p, 32);
  if (counters) { counters[0] = p.path_nodes; counters[1] = p.new_nodes; counters[2] = collectives; }
  return (int)p.error;
}
