// ORACLE — TEST INFRASTRUCTURE ONLY.  C entry points (ctypes) over alpharat_oracle.cpp.
// Loaded by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs only.
#include <atomic>
#include <chrono>
#include <cstring>
#include <thread>
#include <vector>

#include "alpharat_oracle.hpp"

namespace orc {
void fill_search_result_public(ar_search_result& o, const SearchResult& r, const MCTSTree& t);
}

extern "C" {

// predict_fn-style evaluator (mcts/bindings.rs:105-118): n states -> 4 arrays.  rc != 0 = error.
typedef int (*orc_eval_cb)(void* user, const ar_game_pod* states, int n, float* policy_p1,
                           float* policy_p2, float* value_p1, float* value_p2);

struct EvalAdapter {
  orc_eval_cb cb;
  void* user;
};

static int adapter_eval(void* u, const orc::GameState* const* states, int n, orc::EvalResult* out) {
  EvalAdapter* a = (EvalAdapter*)u;
  std::vector<ar_game_pod> pods(n);
  for (int i = 0; i < n; ++i) states[i]->to_pod(pods[i]);
  std::vector<float> p1(n * 5), p2(n * 5), v1(n), v2(n);
  int rc = a->cb(a->user, pods.data(), n, p1.data(), p2.data(), v1.data(), v2.data());
  if (rc != 0) return rc;
  for (int i = 0; i < n; ++i) {
    std::memcpy(out[i].policy_p1, &p1[i * 5], 20);
    std::memcpy(out[i].policy_p2, &p2[i * 5], 20);
    out[i].value_p1 = v1[i];
    out[i].value_p2 = v2[i];
  }
  return 0;
}

// Constant-value evaluator: smart-uniform priors + fixed values (ConstantValueBackend,
// backend.rs:114-129).  user = float[2].
static int const_value_eval(void* u, const orc::GameState* const* states, int n,
                            orc::EvalResult* out) {
  orc::smart_uniform_eval(nullptr, states, n, out);
  const float* v = (const float*)u;
  for (int i = 0; i < n; ++i) {
    out[i].value_p1 = v[0];
    out[i].value_p2 = v[1];
  }
  return 0;
}

// --- rand hooks -------------------------------------------------------------------------
void orc_rng_seed(uint64_t seed, uint64_t state[4]) {
  orc::SmallRng r = orc::SmallRng::seed_from_u64(seed);
  std::memcpy(state, r.s, 32);
}
uint64_t orc_rng_next_u64(uint64_t state[4]) {
  orc::SmallRng r = orc::SmallRng::from_state(state[0], state[1], state[2], state[3]);
  uint64_t v = r.next_u64();
  std::memcpy(state, r.s, 32);
  return v;
}
uint32_t orc_rng_gen_range(uint64_t state[4], uint32_t n) {
  orc::SmallRng r = orc::SmallRng::from_state(state[0], state[1], state[2], state[3]);
  uint32_t v = r.gen_range_u32(n);
  std::memcpy(state, r.s, 32);
  return v;
}
float orc_rng_uniform_f32(uint64_t state[4], float low, float high) {
  orc::SmallRng r = orc::SmallRng::from_state(state[0], state[1], state[2], state[3]);
  float v = r.uniform_f32(low, high);
  std::memcpy(state, r.s, 32);
  return v;
}
uint32_t orc_sample_action(uint64_t state[4], const float policy[5]) {
  orc::SmallRng r = orc::SmallRng::from_state(state[0], state[1], state[2], state[3]);
  uint32_t v = orc::sample_action(policy, r);
  std::memcpy(state, r.s, 32);
  return v;
}

// --- node/tree/search helpers ---------------------------------------------------------------
void orc_compute_outcomes(const uint8_t eff[5], uint8_t outcomes[5], uint8_t* n, uint8_t a2i[5]) {
  orc::compute_outcomes(eff, outcomes, *n, a2i);
}
void orc_smart_uniform_prior(const uint8_t eff[5], float out[5]) {
  orc::smart_uniform_prior(eff, out);
}
void orc_reduce_prior(const float prior5[5], const uint8_t eff[5], float out[5], uint8_t* n) {
  orc::HalfNode h = orc::HalfNode::make(prior5, eff);
  std::memcpy(out, h.prior, 20);
  *n = h.n_outcomes;
}
uint32_t orc_collisions_left(uint32_t node_count, const ar_search_cfg* cfg) {
  return orc::calculate_collisions_left(node_count, orc::SearchConfig::from_c(*cfg));
}
void orc_compute_pruned_visits(const float* q_norm, const float* prior, const float* visits, int n,
                               uint32_t parent_visits, float c_puct, float out[5]) {
  orc::compute_pruned_visits(q_norm, prior, visits, n, parent_visits, c_puct, out);
}

// --- game engine ----------------------------------------------------------------------------
void orc_game_make_move(ar_game_pod* pod, uint8_t d1, uint8_t d2) {
  orc::GameState g = orc::GameState::from_pod(*pod);
  g.make_move(d1, d2);
  g.to_pod(*pod);
}
// make_move then unmake_move; returns 1 when the state is restored bit-for-bit
int orc_game_make_unmake_roundtrip(const ar_game_pod* pod, uint8_t d1, uint8_t d2) {
  orc::GameState g = orc::GameState::from_pod(*pod);
  orc::MoveUndo u = g.make_move(d1, d2);
  g.unmake_move(u);
  ar_game_pod back;
  g.to_pod(back);
  ar_game_pod norm;
  orc::GameState::from_pod(*pod).to_pod(norm);
  return std::memcmp(&back, &norm, sizeof(back)) == 0;
}
void orc_game_effective_actions(const ar_game_pod* pod, int player, uint8_t out[5]) {
  orc::GameState g = orc::GameState::from_pod(*pod);
  if (player == 1) g.effective_actions_p1(out); else g.effective_actions_p2(out);
}
int orc_game_over(const ar_game_pod* pod) {
  return orc::GameState::from_pod(*pod).check_game_over() ? 1 : 0;
}
int orc_obs_dim(int w, int h) { return orc::obs_dim(w, h); }
void orc_encode(const ar_game_pod* pods, int n, float* out) {
  for (int i = 0; i < n; ++i) {
    orc::GameState g = orc::GameState::from_pod(pods[i]);
    orc::encode_flat(g, out + (size_t)i * orc::obs_dim(g.width, g.height));
  }
}

// --- single search: rust_mcts_search (mcts/bindings.rs:228-304) ------------------------------
// eval == NULL -> SmartUniformBackend; const_values != NULL -> ConstantValueBackend.
// clean_out: 1 when every n_in_flight is zero afterwards (search.rs:2750-2791).
int orc_search(const ar_game_pod* pod, const ar_search_cfg* cfg, uint64_t seed, orc_eval_cb eval,
               void* user, const float* const_values, ar_search_result* out, int* clean_out) {
  orc::GameState g = orc::GameState::from_pod(*pod);
  orc::SearchConfig sc = orc::SearchConfig::from_c(*cfg);
  orc::SmallRng rng = orc::SmallRng::seed_from_u64(seed);
  orc::NodeArena* arena = orc::arena_new();
  int rc;
  {
    orc::MCTSTree tree(g, arena);
    orc::SearchResult r;
    EvalAdapter ad{eval, user};
    float cv[2] = {0, 0};
    if (const_values) { cv[0] = const_values[0]; cv[1] = const_values[1]; }
    if (eval)
      rc = orc::run_search(tree, g, adapter_eval, &ad, sc, cfg->simulations, cfg->batch_size, rng, r, nullptr);
    else if (const_values)
      rc = orc::run_search(tree, g, const_value_eval, cv, sc, cfg->simulations, cfg->batch_size, rng, r, nullptr);
    else
      rc = orc::run_search(tree, g, orc::smart_uniform_eval, nullptr, sc, cfg->simulations, cfg->batch_size, rng, r, nullptr);
    if (rc == 0) orc::fill_search_result_public(*out, r, tree);
    if (clean_out) *clean_out = orc::tree_all_in_flight_zero(tree.root) ? 1 : 0;
  }
  orc::arena_free(arena);
  return rc;
}

// --- self-play: run_self_play (selfplay.rs:657-703) with one RNG per game --------------------
// The reference seeds one entropy RNG per worker; for reproducibility each game here gets
// SmallRng::seed_from_u64(seeds[i]) (the same convention as ar_selfplay_run).
int orc_selfplay(const ar_game_pod* pods, int n, const ar_search_cfg* cfg, const uint64_t* seeds,
                 int n_threads, orc_eval_cb eval, void* user, ar_game_summary* summaries,
                 ar_position_record* positions, int stride, ar_stats* stats) {
  orc::SearchConfig sc = orc::SearchConfig::from_c(*cfg);
  std::atomic<int> next{0};
  std::atomic<int> err{0};
  std::atomic<uint64_t> path_nodes{0}, new_nodes{0};
  auto t0 = std::chrono::steady_clock::now();
  auto worker = [&]() {  // game_worker_loop, selfplay.rs:609-650
    orc::NodeArena* arena = orc::arena_new();
    EvalAdapter ad{eval, user};
    orc::SearchCounters ctr;
    for (;;) {
      int idx = next.fetch_add(1);
      if (idx >= n || err.load() != 0) break;
      orc::SmallRng rng = orc::SmallRng::seed_from_u64(seeds[idx]);
      int rc = orc::play_game(orc::GameState::from_pod(pods[idx]),
                              eval ? adapter_eval : orc::smart_uniform_eval,
                              eval ? (void*)&ad : nullptr, sc, cfg->simulations, cfg->batch_size,
                              rng, (uint32_t)idx, arena, summaries[idx],
                              positions + (size_t)idx * stride, stride, &ctr);
      if (rc != 0) err.store(rc);
    }
    path_nodes += ctr.path_nodes;
    new_nodes += ctr.new_nodes;
    orc::arena_free(arena);
  };
  if (n_threads <= 1) {
    worker();
  } else {
    std::vector<std::thread> th;
    for (int i = 0; i < n_threads; ++i) th.emplace_back(worker);
    for (auto& t : th) t.join();
  }
  double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (err.load() != 0) return err.load();
  if (stats) {  // SelfPlayStats::from_games, selfplay.rs:212-224
    std::memset(stats, 0, sizeof(*stats));
    stats->min_turns = 0xFFFFFFFFu;
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
    stats->elapsed_secs = secs;
    stats->path_nodes = path_nodes.load();
    stats->new_nodes = new_nodes.load();
  }
  return 0;
}

uint32_t orc_abi_version(void) { return AR_ABI_VERSION; }

}  // extern "C"
