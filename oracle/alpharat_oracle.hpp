// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement (C++17) of the reference's batched self-play MCTS path.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may build,
// load or call anything in this directory; the product (alpharat_b200/, libalpharat_cuda.so)
// never does.
//
// What it restates (paths relative to the reference tree mintiti/alpharat):
//   node.rs   crates/alpharat-mcts/src/node.rs:57-121,131-283,289-458
//   tree.rs   crates/alpharat-mcts/src/tree.rs:52-94,107-201,209-226,248-302,351-365
//   search.rs crates/alpharat-mcts/src/search.rs:120-152,249-296,362-450,463-554,576-1177
//   backend   crates/alpharat-mcts/src/backend.rs:57-103
//   selfplay  crates/alpharat-sampling/src/selfplay.rs:374-598,609-650
//   encoder   crates/alpharat-sampling/src/flat_encoder.rs:52-124
//
// Third-party arithmetic that is NOT in /root/reference and is restated from the published
// algorithm (PARITY UNPINNED at those boundaries — no reference test fixes their outputs):
//   pyrat-rust 0.2.0 @ 8d10747991f7471ab80778384a07ba54c49dac47 (Cargo.lock:1063-1073) — game step
//   rand 0.8.5 (Cargo.lock:1124-1125)   — SmallRng = xoshiro256++, seed_from_u64, gen_range,
//                                         Uniform<f32>, WeightedIndex<f32>
//   rand_distr 0.4.3 (Cargo.lock:1171-1172) — Gamma for Dirichlet noise: restated as
//                                         Marsaglia–Tsang over a polar-method normal, so
//                                         noise-on runs are distributional, not bit-exact.
// What IS pinned: every golden vector / KAT the reference's own tests hold for this path
// (tests/test_oracle_*.py, oracle/kat_tests.cpp): the 7 encoder fixtures, compute_outcomes
// tables, smart_uniform_prior, calculate_collisions_left integers, backup / Welford KATs,
// pruning cases, reward cases, accounting identities, xoshiro256++ reference vector; the termination rule
// against the reference's own Python statement of it (alpharat/eval/game.py:31-44).
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include "../include/alpharat_cuda.h"

namespace orc {

// ---------------------------------------------------------------------------------------
// rand 0.8.5 restatement
// ---------------------------------------------------------------------------------------
struct SmallRng {
  uint64_t s[4];
  static SmallRng seed_from_u64(uint64_t state);
  static SmallRng from_state(uint64_t a, uint64_t b, uint64_t c, uint64_t d) {
    SmallRng r;
    r.s[0] = a; r.s[1] = b; r.s[2] = c; r.s[3] = d;
    return r;
  }
  uint64_t next_u64();
  uint32_t next_u32() { return (uint32_t)(next_u64() >> 32); }
  // gen_range(0..n) for u32 (UniformInt::sample_single, widening multiply + rejection zone)
  uint32_t gen_range_u32(uint32_t n);
  // Uniform<f32>::new(0, total).sample
  float uniform_f32(float low, float high);
};

// WeightedIndex<f32>::new(policy).sample, STAY(4) fallback on error (selfplay.rs:474-479)
uint8_t sample_action(const float policy[5], SmallRng& rng);

// ---------------------------------------------------------------------------------------
// pyrat-rust GameState restatement
// ---------------------------------------------------------------------------------------
struct MoveUndo {
  uint8_t p1x, p1y, p2x, p2y, mud1, mud2;
  float s1, s2;
  uint16_t turn;
  int n_collected;
  uint16_t collected[2];
};

struct MazeData {  // immutable per game: shared by every clone of the state (cheap clone())
  uint8_t move_cost[AR_MAX_CELLS * 4];
};

struct GameState {
  uint8_t width = 0, height = 0;
  uint16_t turn = 0, max_turns = 0;
  uint8_t p1x = 0, p1y = 0, p2x = 0, p2y = 0, mud1 = 0, mud2 = 0;
  float s1 = 0.f, s2 = 0.f;
  std::shared_ptr<MazeData> maze;
  uint64_t cheese_bits[AR_MAX_CELLS / 64] = {0, 0, 0, 0};
  uint16_t remaining = 0;

  static GameState from_pod(const ar_game_pod& p);
  void to_pod(ar_game_pod& p) const;
  int cell(int x, int y) const { return y * width + x; }
  uint8_t cost(int c, int d) const { return maze->move_cost[c * 4 + d]; }
  bool has_cheese(int c) const { return (cheese_bits[c >> 6] >> (c & 63)) & 1; }
  void set_cheese(int c, bool v) {
    if (v) cheese_bits[c >> 6] |= 1ULL << (c & 63); else cheese_bits[c >> 6] &= ~(1ULL << (c & 63));
  }
  void effective_actions(int x, int y, uint8_t mud, uint8_t out[5]) const;
  void effective_actions_p1(uint8_t out[5]) const { effective_actions(p1x, p1y, mud1, out); }
  void effective_actions_p2(uint8_t out[5]) const { effective_actions(p2x, p2y, mud2, out); }
  MoveUndo make_move(uint8_t d1, uint8_t d2);
  void unmake_move(const MoveUndo& u);
  bool check_game_over() const;
  uint16_t remaining_cheese() const { return remaining; }
};

// ---------------------------------------------------------------------------------------
// node.rs
// ---------------------------------------------------------------------------------------
struct HalfEdge {
  float q = 0.f;
  uint32_t visits = 0;
  uint32_t n_in_flight = 0;
  void update(float value);
  void update_multivisit(float value, uint32_t count);
  uint32_t n_started() const { return visits + n_in_flight; }
};

void compute_outcomes(const uint8_t effective[5], uint8_t outcomes[5], uint8_t& n,
                      uint8_t action_to_idx[5]);

struct HalfNode {
  float prior[5] = {0, 0, 0, 0, 0};
  HalfEdge edges[5];
  uint8_t outcomes[5] = {0, 0, 0, 0, 0};
  uint8_t action_to_idx[5] = {0, 0, 0, 0, 0};
  uint8_t n_outcomes = 0;
  static HalfNode new_shell(const uint8_t effective[5]);
  static HalfNode make(const float prior5[5], const uint8_t effective[5]);
  void set_prior(const float prior5[5]);
  void expand_visits(float out[5]) const;
  void expand_prior(float out[5]) const;
};

struct Node {
  HalfNode p1, p2;
  float v1 = 0.f, v2 = 0.f;
  uint32_t total_visits = 0, n_in_flight = 0;
  float value_scale = 0.f, edge_r1 = 0.f, edge_r2 = 0.f;
  Node* first_child = nullptr;
  Node* next_sibling = nullptr;
  Node* parent = nullptr;
  uint8_t po1 = 0, po2 = 0;
  bool is_terminal = false;
  uint32_t children_visits() const { return total_visits > 0 ? total_visits - 1 : 0; }
  bool try_start_score_update();
  void update_value(float q1, float q2);
  void finalize_score_update(float q1, float q2, uint32_t mv);
};

// ---------------------------------------------------------------------------------------
// tree.rs
// ---------------------------------------------------------------------------------------
struct NodeArena;  // per-thread free list (the reference uses Box + a GC thread)

void smart_uniform_prior(const uint8_t effective[5], float out[5]);

struct EvalResult {
  float policy_p1[5], policy_p2[5];
  float value_p1, value_p2;
};

struct MCTSTree {
  Node* root = nullptr;
  uint32_t node_count = 0;
  NodeArena* arena = nullptr;
  explicit MCTSTree(const GameState& g, NodeArena* a);
  ~MCTSTree();
  MCTSTree(const MCTSTree&) = delete;
  bool advance_root(uint8_t a1, uint8_t a2);
  void reinit(const GameState& g);
};

Node* find_child(Node* parent, uint8_t i, uint8_t j);
Node* extend_node(NodeArena* arena, Node* parent, uint8_t i, uint8_t j, const GameState& g);
void populate_node(Node* node, const EvalResult* r);

// ---------------------------------------------------------------------------------------
// search.rs
// ---------------------------------------------------------------------------------------
struct SearchConfig {
  float c_puct = 1.5f, fpu_reduction = 0.2f, force_k = 2.0f;
  float noise_epsilon = 0.f, noise_concentration = 10.83f;
  uint32_t collision_limit_min = 1, collision_limit_max = 256;
  uint32_t collision_scaling_start = 800, collision_scaling_end = 50000;
  float collision_scaling_power = 1.0f;
  static SearchConfig from_c(const ar_search_cfg& c);
};

// evaluate_batch: returns 0 on success.  states[i] -> out[i].
typedef int (*EvalFn)(void* user, const GameState* const* states, int n, EvalResult* out);
int smart_uniform_eval(void*, const GameState* const* states, int n, EvalResult* out);

struct SearchResult {
  float policy_p1[5], policy_p2[5];
  float value_p1, value_p2;
  float visit_counts_p1[5], visit_counts_p2[5];
  float prior_p1[5], prior_p2[5];
  uint32_t total_visits, nn_evals, terminals, collisions;
};

struct SearchCounters {  // roofline accounting; not part of the reference
  uint64_t path_nodes = 0, new_nodes = 0;
};

uint32_t calculate_collisions_left(uint32_t tree_node_count, const SearchConfig& c);
void compute_pruned_visits(const float* q_norm, const float* prior, const float* visits, int n,
                           uint32_t parent_visits, float c_puct, float out[5]);
void backup(const std::vector<Node*>& path_nodes, const std::vector<uint8_t>& a1,
            const std::vector<uint8_t>& a2, Node* leaf, float g1, float g2);
void backup_and_finalize(Node* leaf, float g1, float g2, uint32_t mv, SearchCounters* ctr);
// returns 0 ok, nonzero = backend error (all virtual losses reverted)
int run_search(MCTSTree& tree, const GameState& game, EvalFn eval, void* user,
               const SearchConfig& cfg, uint32_t n_sims, uint32_t batch_size, SmallRng& rng,
               SearchResult& out, SearchCounters* ctr);
bool tree_all_in_flight_zero(const Node* root);

// ---------------------------------------------------------------------------------------
// selfplay.rs
// ---------------------------------------------------------------------------------------
int play_game(GameState game, EvalFn eval, void* user, const SearchConfig& cfg, uint32_t n_sims,
              uint32_t batch_size, SmallRng& rng, uint32_t game_index, NodeArena* arena,
              ar_game_summary& summary, ar_position_record* positions, int positions_cap,
              SearchCounters* ctr);

// flat_encoder.rs
int obs_dim(int w, int h);
void encode_flat(const GameState& g, float* out);

NodeArena* arena_new();
void arena_free(NodeArena*);

}  // namespace orc
