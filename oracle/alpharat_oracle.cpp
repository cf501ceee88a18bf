// ORACLE — TEST INFRASTRUCTURE ONLY (see alpharat_oracle.hpp for scope, citations and the
// "parity unpinned" statement for the third-party pieces).
//
// Build: g++ -O2 -std=c++17 -ffp-contract=off -fno-fast-math (oracle/Makefile).  All float
// arithmetic is plain IEEE f32 in the reference's operation order; no FMA contraction.
#include "alpharat_oracle.hpp"

#include <algorithm>
#include <cassert>
#include <cmath>
#include <limits>

namespace orc {

// =======================================================================================
// rand 0.8.5 — SmallRng (xoshiro256++), restated from the published algorithm
// =======================================================================================
static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

SmallRng SmallRng::seed_from_u64(uint64_t state) {
  // SplitMix64 fills the 32-byte seed, little-endian (rand_xoshiro / rand 0.8.5 SeedableRng)
  SmallRng r;
  for (int i = 0; i < 4; ++i) {
    state += 0x9e3779b97f4a7c15ULL;
    uint64_t z = state;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    z = z ^ (z >> 31);
    r.s[i] = z;
  }
  return r;
}

uint64_t SmallRng::next_u64() {
  uint64_t result = rotl64(s[0] + s[3], 23) + s[0];
  uint64_t t = s[1] << 17;
  s[2] ^= s[0];
  s[3] ^= s[1];
  s[1] ^= s[2];
  s[0] ^= s[3];
  s[2] ^= t;
  s[3] = rotl64(s[3], 45);
  return result;
}

uint32_t SmallRng::gen_range_u32(uint32_t n) {
  // UniformInt<u32>::sample_single(0, n) -> sample_single_inclusive(0, n-1)
  uint32_t range = n;  // high - low + 1
  if (range == 0) return next_u32();
  int lz = __builtin_clz(range);
  uint32_t zone = (range << lz) - 1u;
  for (;;) {
    uint32_t v = next_u32();
    uint64_t m = (uint64_t)v * (uint64_t)range;
    uint32_t hi = (uint32_t)(m >> 32), lo = (uint32_t)m;
    if (lo <= zone) return hi;
  }
}

static inline float f32_from_bits(uint32_t b) {
  float f;
  std::memcpy(&f, &b, 4);
  return f;
}
static inline uint32_t f32_bits(float f) {
  uint32_t b;
  std::memcpy(&b, &f, 4);
  return b;
}

float SmallRng::uniform_f32(float low, float high) {
  // UniformFloat<f32>::new(low, high) then sample()
  const float max_rand = f32_from_bits((0xFFFFFFFFu >> 9) | (127u << 23)) - 1.0f;
  float scale = high - low;
  for (;;) {
    float top = scale * max_rand + low;
    if (!(top >= high)) break;
    scale = f32_from_bits(f32_bits(scale) - 1);
  }
  float value1_2 = f32_from_bits((next_u32() >> 9) | (127u << 23));
  float value0_1 = value1_2 - 1.0f;
  return value0_1 * scale + low;
}

uint8_t sample_action(const float policy[5], SmallRng& rng) {
  // WeightedIndex::new: first weight, then cumulative (excluding the last)
  float total = policy[0];
  if (!(total >= 0.0f)) return 4;
  float cum[4];
  for (int i = 1; i < 5; ++i) {
    if (!(policy[i] >= 0.0f)) return 4;
    cum[i - 1] = total;
    total += policy[i];
  }
  if (total == 0.0f) return 4;
  if (!std::isfinite(total)) return 4;  // Uniform::new would panic; unreachable for policies
  float x = rng.uniform_f32(0.0f, total);
  // binary_search_by(|w| if *w <= x {Less} else {Greater}).unwrap_err() == #weights <= x
  uint8_t idx = 0;
  while (idx < 4 && cum[idx] <= x) ++idx;
  return idx;
}

// =======================================================================================
// pyrat-rust GameState (restated; Appendix B.1 of SURVEY.md)
// =======================================================================================
static const int DX[5] = {0, 1, 0, -1, 0};
static const int DY[5] = {1, 0, -1, 0, 0};

GameState GameState::from_pod(const ar_game_pod& p) {
  GameState g;
  g.width = p.width; g.height = p.height;
  g.turn = p.turn; g.max_turns = p.max_turns;
  g.p1x = p.p1_x; g.p1y = p.p1_y; g.p2x = p.p2_x; g.p2y = p.p2_y;
  g.mud1 = p.p1_mud; g.mud2 = p.p2_mud;
  g.s1 = p.p1_score; g.s2 = p.p2_score;
  auto m = std::make_shared<MazeData>();
  std::memcpy(m->move_cost, p.move_cost, sizeof(m->move_cost));
  g.maze = m;
  g.remaining = 0;
  int cells = (int)p.width * p.height;
  for (int c = 0; c < cells; ++c)
    if ((p.cheese[c >> 3] >> (c & 7)) & 1) {
      g.set_cheese(c, true);
      g.remaining++;
    }
  return g;
}

void GameState::to_pod(ar_game_pod& p) const {
  std::memset(&p, 0, sizeof(p));
  p.width = width; p.height = height; p.turn = turn; p.max_turns = max_turns;
  p.p1_x = p1x; p.p1_y = p1y; p.p2_x = p2x; p.p2_y = p2y;
  p.p1_mud = mud1; p.p2_mud = mud2; p.p1_score = s1; p.p2_score = s2;
  std::memcpy(p.move_cost, maze->move_cost, sizeof(maze->move_cost));
  for (int c = 0; c < AR_MAX_CELLS; ++c)
    if (has_cheese(c)) p.cheese[c >> 3] |= (uint8_t)(1u << (c & 7));
}

void GameState::effective_actions(int x, int y, uint8_t mud, uint8_t out[5]) const {
  // game.pyi:337-376: blocked -> STAY(4); a player in mud has [4,4,4,4,4]
  for (int a = 0; a < 5; ++a) out[a] = 4;
  if (mud > 0) return;
  int c = cell(x, y);
  for (int a = 0; a < 4; ++a)
    if (cost(c, a) != 0) out[a] = (uint8_t)a;
}

static void step_player(const GameState& g, uint8_t& x, uint8_t& y, uint8_t& mud, uint8_t d) {
  if (mud > 0) {  // stuck: the timer runs down, the move is ignored
    mud -= 1;
    return;
  }
  if (d >= 4) return;
  uint8_t cost = g.cost(g.cell(x, y), d);
  if (cost == 0) return;  // wall / boundary: a blocked move is a stay
  x = (uint8_t)(x + DX[d]);
  y = (uint8_t)(y + DY[d]);
  if (cost >= 2) mud = cost;  // fixtures/mud_stuck_5x5.json: position = target, timer = cost
}

MoveUndo GameState::make_move(uint8_t d1, uint8_t d2) {
  MoveUndo u;
  u.p1x = p1x; u.p1y = p1y; u.p2x = p2x; u.p2y = p2y; u.mud1 = mud1; u.mud2 = mud2;
  u.s1 = s1; u.s2 = s2; u.turn = turn; u.n_collected = 0;
  step_player(*this, p1x, p1y, mud1, d1);
  step_player(*this, p2x, p2y, mud2, d2);
  // cheese collection: only players not stuck in mud collect (tree.rs:934-999 pins the amounts)
  bool c1 = mud1 == 0, c2 = mud2 == 0;
  int k1 = cell(p1x, p1y), k2 = cell(p2x, p2y);
  if (c1 && c2 && k1 == k2) {
    if (has_cheese(k1)) {
      set_cheese(k1, false); remaining--;
      s1 += 0.5f; s2 += 0.5f;
      u.collected[u.n_collected++] = (uint16_t)k1;
    }
  } else {
    if (c1 && has_cheese(k1)) {
      set_cheese(k1, false); remaining--;
      s1 += 1.0f;
      u.collected[u.n_collected++] = (uint16_t)k1;
    }
    if (c2 && has_cheese(k2)) {
      set_cheese(k2, false); remaining--;
      s2 += 1.0f;
      u.collected[u.n_collected++] = (uint16_t)k2;
    }
  }
  turn += 1;
  return u;
}

void GameState::unmake_move(const MoveUndo& u) {
  p1x = u.p1x; p1y = u.p1y; p2x = u.p2x; p2y = u.p2y; mud1 = u.mud1; mud2 = u.mud2;
  s1 = u.s1; s2 = u.s2; turn = u.turn;
  for (int i = 0; i < u.n_collected; ++i) {
    set_cheese(u.collected[i], true);
    remaining++;
  }
}

bool GameState::check_game_over() const {
  // alpharat/eval/game.py:31-44
  if (turn >= max_turns) return true;
  if (remaining == 0) return true;
  float total = s1 + s2 + (float)remaining;
  return s1 > total / 2.0f || s2 > total / 2.0f;
}

// =======================================================================================
// node.rs
// =======================================================================================
void HalfEdge::update(float value) {  // node.rs:75-78
  visits += 1;
  q += (value - q) / (float)visits;
}
void HalfEdge::update_multivisit(float value, uint32_t count) {  // node.rs:82-85
  visits += count;
  q += (value - q) * (float)count / (float)visits;
}

void compute_outcomes(const uint8_t effective[5], uint8_t unique[5], uint8_t& n,
                      uint8_t action_to_idx[5]) {  // node.rs:251-283
  for (int i = 0; i < 5; ++i) unique[i] = 0;
  n = 0;
  for (int a = 0; a < 5; ++a) {
    uint8_t val = effective[a];
    int pos = 0;
    while (pos < n && unique[pos] < val) ++pos;
    if (pos < n && unique[pos] == val) continue;
    for (int i = n; i > pos; --i) unique[i] = unique[i - 1];
    unique[pos] = val;
    n++;
  }
  for (int a = 0; a < 5; ++a) {
    int idx = 0;
    while (idx < n && unique[idx] < effective[a]) ++idx;
    action_to_idx[a] = (uint8_t)idx;
  }
}

HalfNode HalfNode::new_shell(const uint8_t effective[5]) {  // node.rs:158-167
  HalfNode h;
  compute_outcomes(effective, h.outcomes, h.n_outcomes, h.action_to_idx);
  return h;
}
HalfNode HalfNode::make(const float prior5[5], const uint8_t effective[5]) {  // node.rs:148-152
  HalfNode h = new_shell(effective);
  h.set_prior(prior5);
  return h;
}
void HalfNode::set_prior(const float prior5[5]) {  // node.rs:173-179
  for (int i = 0; i < 5; ++i) prior[i] = 0.f;
  for (int a = 0; a < 5; ++a) prior[action_to_idx[a]] += prior5[a];
}
void HalfNode::expand_visits(float out[5]) const {  // node.rs:216-223
  for (int i = 0; i < 5; ++i) out[i] = 0.f;
  for (int i = 0; i < n_outcomes; ++i) out[outcomes[i]] = (float)edges[i].visits;
}
void HalfNode::expand_prior(float out[5]) const {  // node.rs:234-240
  for (int i = 0; i < 5; ++i) out[i] = 0.f;
  for (int i = 0; i < n_outcomes; ++i) out[outcomes[i]] = prior[i];
}

bool Node::try_start_score_update() {  // node.rs:388-394
  if (total_visits == 0 && n_in_flight > 0) return false;
  n_in_flight += 1;
  return true;
}
void Node::update_value(float q1, float q2) {  // node.rs:435-440
  total_visits += 1;
  float n = (float)total_visits;
  v1 += (q1 - v1) / n;
  v2 += (q2 - v2) / n;
}
void Node::finalize_score_update(float q1, float q2, uint32_t mv) {  // node.rs:444-457
  total_visits += mv;
  float n = (float)total_visits;
  float w = (float)mv;
  v1 += (q1 - v1) * w / n;
  v2 += (q2 - v2) * w / n;
  assert(n_in_flight >= mv);
  n_in_flight -= mv;
}

// =======================================================================================
// tree.rs
// =======================================================================================
struct NodeArena {
  std::vector<Node*> free_list;
  std::vector<Node*> blocks;
  Node* alloc() {
    if (free_list.empty()) {
      const int B = 1024;
      Node* blk = new Node[B];
      blocks.push_back(blk);
      for (int i = B - 1; i >= 0; --i) free_list.push_back(blk + i);
    }
    Node* n = free_list.back();
    free_list.pop_back();
    *n = Node();
    return n;
  }
  void release_subtree(Node* root) {  // iterative, like Node::drop (node.rs:464-484)
    std::vector<Node*> stack;
    stack.push_back(root);
    while (!stack.empty()) {
      Node* n = stack.back();
      stack.pop_back();
      if (n->first_child) stack.push_back(n->first_child);
      if (n->next_sibling) stack.push_back(n->next_sibling);
      free_list.push_back(n);
    }
  }
  ~NodeArena() {
    for (Node* b : blocks) delete[] b;
  }
};
NodeArena* arena_new() { return new NodeArena(); }
void arena_free(NodeArena* a) { delete a; }

void smart_uniform_prior(const uint8_t effective[5], float prior[5]) {  // tree.rs:69-84
  bool seen[5] = {false, false, false, false, false};
  int count = 0;
  for (int a = 0; a < 5; ++a)
    if (!seen[effective[a]]) {
      seen[effective[a]] = true;
      count++;
    }
  float p = 1.0f / (float)count;
  for (int i = 0; i < 5; ++i) prior[i] = 0.f;
  for (int a = 0; a < 5; ++a) prior[effective[a]] = p;
}

Node* find_child(Node* parent, uint8_t i, uint8_t j) {  // tree.rs:52-63
  for (Node* c = parent->first_child; c; c = c->next_sibling)
    if (c->po1 == i && c->po2 == j) return c;
  return nullptr;
}

Node* extend_node(NodeArena* arena, Node* parent, uint8_t i, uint8_t j,
                  const GameState& g) {  // tree.rs:107-148
  uint8_t e1[5], e2[5];
  g.effective_actions_p1(e1);
  g.effective_actions_p2(e2);
  Node* n = arena->alloc();
  n->p1 = HalfNode::new_shell(e1);
  n->p2 = HalfNode::new_shell(e2);
  n->value_scale = (float)std::max<uint16_t>(g.remaining_cheese(), 1);
  n->parent = parent;
  n->po1 = i;
  n->po2 = j;
  n->next_sibling = parent->first_child;  // prepend
  parent->first_child = n;
  return n;
}

void populate_node(Node* node, const EvalResult* r) {  // tree.rs:156-173
  assert(node->total_visits == 0);
  if (r) {
    node->p1.set_prior(r->policy_p1);
    node->p2.set_prior(r->policy_p2);
  } else {
    node->is_terminal = true;
  }
}

static Node* alloc_root(NodeArena* arena, const GameState& g) {  // tree.rs:351-365
  uint8_t e1[5], e2[5];
  g.effective_actions_p1(e1);
  g.effective_actions_p2(e2);
  float pr1[5], pr2[5];
  smart_uniform_prior(e1, pr1);
  smart_uniform_prior(e2, pr2);
  Node* n = arena->alloc();
  n->p1 = HalfNode::make(pr1, e1);
  n->p2 = HalfNode::make(pr2, e2);
  n->value_scale = (float)std::max<uint16_t>(g.remaining_cheese(), 1);
  return n;
}

static uint32_t count_subtree_nodes(const Node* root) {  // tree.rs:209-226
  uint32_t count = 1;
  std::vector<const Node*> stack;
  if (root->first_child) stack.push_back(root->first_child);
  while (!stack.empty()) {
    const Node* n = stack.back();
    stack.pop_back();
    count++;
    if (n->first_child) stack.push_back(n->first_child);
    if (n->next_sibling) stack.push_back(n->next_sibling);
  }
  return count;
}

MCTSTree::MCTSTree(const GameState& g, NodeArena* a) : arena(a) {  // tree.rs:248-255
  root = alloc_root(arena, g);
  node_count = 1;
}
MCTSTree::~MCTSTree() {
  if (root) arena->release_subtree(root);
}

bool MCTSTree::advance_root(uint8_t a1, uint8_t a2) {  // tree.rs:283-339
  uint8_t i = root->p1.action_to_idx[a1];
  uint8_t j = root->p2.action_to_idx[a2];
  Node* prev = nullptr;
  Node* c = root->first_child;
  while (c && !(c->po1 == i && c->po2 == j)) {
    prev = c;
    c = c->next_sibling;
  }
  if (!c) return false;
  if (prev) prev->next_sibling = c->next_sibling; else root->first_child = c->next_sibling;
  c->next_sibling = nullptr;
  c->parent = nullptr;
  arena->release_subtree(root);
  root = c;
  node_count = count_subtree_nodes(root);
  return true;
}

void MCTSTree::reinit(const GameState& g) {  // tree.rs:298-302
  arena->release_subtree(root);
  root = alloc_root(arena, g);
  node_count = 1;
}

// =======================================================================================
// search.rs
// =======================================================================================
static const float FORCED_PLAYOUT_SCORE = 1e20f;
static const float NEG_INF = -std::numeric_limits<float>::infinity();

SearchConfig SearchConfig::from_c(const ar_search_cfg& c) {
  SearchConfig s;
  s.c_puct = c.c_puct; s.fpu_reduction = c.fpu_reduction; s.force_k = c.force_k;
  s.noise_epsilon = c.noise_epsilon; s.noise_concentration = c.noise_concentration;
  s.collision_limit_min = c.collision_limit_min; s.collision_limit_max = c.collision_limit_max;
  s.collision_scaling_start = c.collision_scaling_start;
  s.collision_scaling_end = c.collision_scaling_end;
  s.collision_scaling_power = c.collision_scaling_power;
  return s;
}

int smart_uniform_eval(void*, const GameState* const* states, int n,
                       EvalResult* out) {  // backend.rs:94-103
  for (int i = 0; i < n; ++i) {
    uint8_t e1[5], e2[5];
    states[i]->effective_actions_p1(e1);
    states[i]->effective_actions_p2(e2);
    smart_uniform_prior(e1, out[i].policy_p1);
    smart_uniform_prior(e2, out[i].policy_p2);
    out[i].value_p1 = 0.f;
    out[i].value_p2 = 0.f;
  }
  return 0;
}

// Legacy single-visit backup (search.rs:75-111); kept for the reference's backup KATs.
void backup(const std::vector<Node*>& path_nodes, const std::vector<uint8_t>& a1,
            const std::vector<uint8_t>& a2, Node* leaf, float g1, float g2) {
  leaf->update_value(g1, g2);
  float v1 = g1, v2 = g2;
  Node* child = leaf;
  for (int k = (int)path_nodes.size() - 1; k >= 0; --k) {
    float q1 = child->edge_r1 + v1;
    float q2 = child->edge_r2 + v2;
    Node* node = path_nodes[k];
    node->update_value(q1, q2);
    node->p1.edges[a1[k]].update(q1);
    node->p2.edges[a2[k]].update(q2);
    v1 = q1; v2 = q2;
    child = node;
  }
}

static float compute_fpu(const HalfNode& half, float node_value, float value_scale,
                         float fpu_reduction) {  // search.rs:120-128
  float mass = 0.f;
  for (int i = 0; i < half.n_outcomes; ++i)
    if (half.edges[i].visits > 0) mass += half.prior[i];
  return node_value - fpu_reduction * value_scale * std::sqrt(mass);
}

static inline void puct_score(const HalfEdge& edge, float prior, float fpu, float value_scale,
                              float c_puct, float sqrt_total, float nstarted, float& score,
                              float& q_norm) {  // search.rs:139-152
  float q = edge.visits > 0 ? edge.q : fpu;
  q_norm = q / value_scale;
  float exploration = c_puct * prior * sqrt_total / (1.0f + nstarted);
  score = q_norm + exploration;
}

void compute_pruned_visits(const float* q_norm, const float* prior, const float* visits, int n,
                           uint32_t parent_visits, float c_puct,
                           float result[5]) {  // search.rs:249-296
  for (int i = 0; i < 5; ++i) result[i] = 0.f;
  if (n <= 1) {
    if (n == 1) result[0] = visits[0];
    return;
  }
  int best_idx = 0;
  float best_visits = visits[0];
  for (int i = 1; i < n; ++i)
    if (visits[i] > best_visits) {
      best_visits = visits[i];
      best_idx = i;
    }
  float sqrt_total = std::sqrt((float)std::max<uint32_t>(parent_visits, 1));
  float puct_star =
      q_norm[best_idx] + c_puct * prior[best_idx] * sqrt_total / (1.0f + visits[best_idx]);
  for (int i = 0; i < n; ++i) {
    if (i == best_idx || q_norm[i] >= puct_star) {
      result[i] = visits[i];
    } else {
      float denom = puct_star - q_norm[i];
      if (denom <= 0.0f) {
        result[i] = visits[i];
      } else {
        float n_min = std::fmax(c_puct * prior[i] * sqrt_total / denom - 1.0f, 0.0f);
        result[i] = std::fmin(visits[i], n_min);
      }
    }
  }
}

uint32_t calculate_collisions_left(uint32_t n, const SearchConfig& c) {  // search.rs:437-450
  if (n >= c.collision_scaling_end) return c.collision_limit_max;
  if (n <= c.collision_scaling_start) return c.collision_limit_min;
  float ratio = (float)(n - c.collision_scaling_start) /
                (float)(c.collision_scaling_end - c.collision_scaling_start);
  float scaled = (float)c.collision_limit_min +
                 ((float)c.collision_limit_max - (float)c.collision_limit_min) *
                     std::pow(ratio, c.collision_scaling_power);
  float r = std::round(scaled);  // half away from zero, like f32::round
  uint32_t v;
  if (!(r > 0.0f)) v = 0; else if (r >= 4294967296.0f) v = 0xFFFFFFFFu; else v = (uint32_t)r;
  return std::min(std::max(v, c.collision_limit_min), c.collision_limit_max);
}

static inline uint32_t f32_to_u32_sat(float f) {  // Rust `as u32`
  if (!(f > 0.0f)) return 0;
  if (f >= 4294967296.0f) return 0xFFFFFFFFu;
  return (uint32_t)f;
}

static void estimated_visits_to_change_best_half(const HalfNode& half, float node_value,
                                                 float value_scale, uint32_t children_visits,
                                                 const SearchConfig& cfg, bool is_root,
                                                 const uint32_t nstarted[5], SmallRng& rng,
                                                 uint8_t& best_out,
                                                 uint32_t& vtc_out) {  // search.rs:463-554
  int n = half.n_outcomes;
  if (n <= 1) {
    best_out = 0;
    vtc_out = 0xFFFFFFFFu;
    return;
  }
  float fpu = compute_fpu(half, node_value, value_scale, cfg.fpu_reduction);
  float sqrt_total = std::sqrt((float)std::max<uint32_t>(children_visits, 1));
  float c_puct = cfg.c_puct;

  uint8_t best_idx = 0;
  float best_score = NEG_INF, best_utility = NEG_INF, second_best_score = NEG_INF;
  float scores[5], qn[5];
  for (int i = 0; i < n; ++i) {
    float score, q_norm;
    puct_score(half.edges[i], half.prior[i], fpu, value_scale, c_puct, sqrt_total,
               (float)nstarted[i], score, q_norm);
    if (is_root && cfg.force_k > 0.0f && half.prior[i] > 0.0f) {
      float threshold = std::sqrt(cfg.force_k * half.prior[i] * (float)children_visits);
      if ((float)half.edges[i].visits < threshold) score = FORCED_PLAYOUT_SCORE;
    }
    scores[i] = score;
    qn[i] = q_norm;
    if (score > best_score) {
      second_best_score = best_score;
      best_score = score;
      best_idx = (uint8_t)i;
      best_utility = q_norm;
    } else if (score > second_best_score) {
      second_best_score = score;
    }
  }
  uint32_t tie_count = 1;
  for (int i = 0; i < n; ++i) {
    if ((uint8_t)i == best_idx) continue;
    if (std::fabs(scores[i] - best_score) < 1e-12f) {
      tie_count += 1;
      if (rng.gen_range_u32(tie_count) == 0) {
        best_idx = (uint8_t)i;
        best_utility = qn[i];
      }
    }
  }
  best_out = best_idx;
  if (second_best_score <= NEG_INF) { vtc_out = 0xFFFFFFFFu; return; }
  if (best_utility >= second_best_score) { vtc_out = 0xFFFFFFFFu; return; }
  float prior_best = half.prior[best_idx];
  float n1 = (float)nstarted[best_idx] + 1.0f;
  float denom = second_best_score - best_utility;
  if (denom <= 0.0f) { vtc_out = 0xFFFFFFFFu; return; }
  float vtc = std::fmax(c_puct * prior_best * sqrt_total / denom - n1 + 1.0f, 1.0f);
  vtc_out = std::max<uint32_t>(f32_to_u32_sat(vtc), 1);
}

struct GatherLevel {
  Node* node;
  uint32_t vtp[25];
  int next_idx, last_idx;
};

static GatherLevel build_gather_level(Node* node, uint32_t cur_limit, const SearchConfig& cfg,
                                      bool is_root, SmallRng& rng) {  // search.rs:742-817
  int n1 = node->p1.n_outcomes, n2 = node->p2.n_outcomes;
  uint32_t children_visits = node->children_visits();
  float value_scale = node->value_scale, v1 = node->v1, v2 = node->v2;
  uint32_t ns1[5] = {0, 0, 0, 0, 0}, ns2[5] = {0, 0, 0, 0, 0};
  for (int i = 0; i < n1; ++i) ns1[i] = node->p1.edges[i].n_started();
  for (int j = 0; j < n2; ++j) ns2[j] = node->p2.edges[j].n_started();
  uint32_t o1[5], o2[5];
  std::memcpy(o1, ns1, sizeof(o1));
  std::memcpy(o2, ns2, sizeof(o2));
  GatherLevel lv;
  lv.node = node;
  std::memset(lv.vtp, 0, sizeof(lv.vtp));
  lv.next_idx = 0;
  lv.last_idx = 0;
  uint32_t remaining = cur_limit;
  while (remaining > 0) {
    uint8_t b1, b2;
    uint32_t t1, t2;
    estimated_visits_to_change_best_half(node->p1, v1, value_scale, children_visits, cfg, is_root,
                                         ns1, rng, b1, t1);
    estimated_visits_to_change_best_half(node->p2, v2, value_scale, children_visits, cfg, is_root,
                                         ns2, rng, b2, t2);
    uint32_t k = std::max<uint32_t>(std::min(std::min(remaining, t1), t2), 1);
    int flat = b1 * 5 + b2;
    lv.vtp[flat] += k;
    ns1[b1] += k;
    ns2[b2] += k;
    remaining -= k;
    if (flat > lv.last_idx && lv.vtp[flat] > 0) lv.last_idx = flat;
  }
  for (int i = 0; i < n1; ++i) node->p1.edges[i].n_in_flight += ns1[i] - o1[i];
  for (int j = 0; j < n2; ++j) node->p2.edges[j].n_in_flight += ns2[j] - o2[j];
  return lv;
}

enum class Kind : uint8_t { NeedsEval, Terminal };
struct NodeToProcess {
  Node* node;
  Kind kind;
  uint32_t multivisit;
  GameState state;  // valid for NeedsEval
};
struct Collision {
  Node* node;
  uint32_t mv;
};

static void pick_nodes_to_extend(MCTSTree& tree, const GameState& game, const SearchConfig& cfg,
                                 uint32_t budget, SmallRng& rng, std::vector<NodeToProcess>& tp,
                                 std::vector<Collision>& coll,
                                 SearchCounters* ctr) {  // search.rs:576-738
  Node* root = tree.root;
  GameState work = game;
  std::vector<MoveUndo> undos;
  uint32_t cur_limit = budget;

  if (root->total_visits == 0 || root->is_terminal) {
    if (root->total_visits == 0 && !root->is_terminal) {
      if (root->try_start_score_update()) {
        if (work.check_game_over()) {
          populate_node(root, nullptr);
          tp.push_back({root, Kind::Terminal, 1, GameState()});
        } else {
          tp.push_back({root, Kind::NeedsEval, 1, work});
        }
        if (cur_limit > 1) coll.push_back({root, cur_limit - 1});
      } else {
        coll.push_back({root, cur_limit});
      }
    } else {
      if (root->total_visits == 0) populate_node(root, nullptr);
      if (root->try_start_score_update()) {
        tp.push_back({root, Kind::Terminal, 1, GameState()});
        if (cur_limit > 1) coll.push_back({root, cur_limit - 1});
      } else {
        coll.push_back({root, cur_limit});
      }
    }
    return;
  }

  root->n_in_flight += cur_limit;
  std::vector<GatherLevel> levels;
  levels.push_back(build_gather_level(root, cur_limit, cfg, true, rng));

  while (!levels.empty()) {
    bool found_child = false;
    while (levels.back().next_idx <= levels.back().last_idx) {
      GatherLevel& level = levels.back();
      int idx = level.next_idx;
      level.next_idx += 1;
      if (level.vtp[idx] == 0) continue;
      uint8_t a1 = (uint8_t)(idx / 5), a2 = (uint8_t)(idx % 5);
      uint32_t k = level.vtp[idx];
      uint8_t act1 = level.node->p1.outcomes[a1], act2 = level.node->p2.outcomes[a2];
      float sb1 = work.s1, sb2 = work.s2;
      MoveUndo undo = work.make_move(act1, act2);
      float r1 = work.s1 - sb1, r2 = work.s2 - sb2;  // compute_rewards, tree.rs:89-94

      Node* child = find_child(level.node, a1, a2);  // find_or_extend_child, tree.rs:186-201
      if (!child) {
        child = extend_node(tree.arena, level.node, a1, a2, work);
        child->edge_r1 = r1;
        child->edge_r2 = r2;
        tree.node_count += 1;
        if (ctr) ctr->new_nodes++;
      }
      if (child->total_visits == 0 || child->is_terminal) {
        if (child->try_start_score_update()) {
          if (child->is_terminal || work.check_game_over()) {
            if (child->total_visits == 0) populate_node(child, nullptr);
            tp.push_back({child, Kind::Terminal, 1, GameState()});
          } else {
            tp.push_back({child, Kind::NeedsEval, 1, work});
          }
          if (k > 1) coll.push_back({child, k - 1});
        } else {
          coll.push_back({child, k});
        }
        work.unmake_move(undo);
      } else {
        // visited interior child: claim always succeeds
        child->try_start_score_update();
        if (k > 1) child->n_in_flight += k - 1;
        undos.push_back(undo);
        GatherLevel cl = build_gather_level(child, k, cfg, false, rng);
        levels.push_back(cl);
        found_child = true;
        break;
      }
    }
    if (!found_child) {
      levels.pop_back();
      if (!undos.empty()) {
        work.unmake_move(undos.back());
        undos.pop_back();
      }
    }
  }
}

void backup_and_finalize(Node* leaf, float g1, float g2, uint32_t mv,
                         SearchCounters* ctr) {  // search.rs:826-852
  leaf->finalize_score_update(g1, g2, mv);
  float v1 = g1, v2 = g2;
  Node* current = leaf;
  while (Node* parent = current->parent) {
    float q1 = current->edge_r1 + v1;
    float q2 = current->edge_r2 + v2;
    parent->finalize_score_update(q1, q2, mv);
    HalfEdge& e1 = parent->p1.edges[current->po1];
    HalfEdge& e2 = parent->p2.edges[current->po2];
    e1.update_multivisit(q1, mv);
    e2.update_multivisit(q2, mv);
    assert(e1.n_in_flight >= mv && e2.n_in_flight >= mv);
    e1.n_in_flight -= mv;
    e2.n_in_flight -= mv;
    v1 = q1; v2 = q2;
    current = parent;
    (void)ctr;
  }
}

static void cancel_shared_collisions(const std::vector<Collision>& coll,
                                     Node* root) {  // search.rs:860-889
  for (const Collision& c : coll) {
    Node* current = c.node;
    while (Node* parent = current->parent) {
      assert(parent->n_in_flight >= c.mv);
      parent->n_in_flight -= c.mv;
      parent->p1.edges[current->po1].n_in_flight -= c.mv;
      parent->p2.edges[current->po2].n_in_flight -= c.mv;
      if (parent == root) break;
      current = parent;
    }
  }
}

static void cancel_leaf_and_path(Node* leaf, uint32_t mv) {  // search.rs:899-910
  leaf->n_in_flight -= mv;
  Node* current = leaf;
  while (Node* parent = current->parent) {
    parent->n_in_flight -= mv;
    parent->p1.edges[current->po1].n_in_flight -= mv;
    parent->p2.edges[current->po2].n_in_flight -= mv;
    current = parent;
  }
}

// --- Dirichlet noise (search.rs:400-429).  rand_distr's Gamma is NOT restated bit-for-bit:
// Marsaglia–Tsang with the alpha<1 boost, over a polar-method normal.  Distributional only.
static double open01_f64(SmallRng& rng) {
  uint64_t bits = (rng.next_u64() >> 12) | (1023ULL << 52);
  double d;
  std::memcpy(&d, &bits, 8);
  return d - (1.0 - 2.220446049250313e-16 / 2.0);
}
static double std_normal(SmallRng& rng) {
  for (;;) {
    double u = 2.0 * open01_f64(rng) - 1.0, v = 2.0 * open01_f64(rng) - 1.0;
    double s = u * u + v * v;
    if (s > 0.0 && s < 1.0) return u * std::sqrt(-2.0 * std::log(s) / s);
  }
}
static double gamma_large(double shape, SmallRng& rng) {
  double d = shape - 1.0 / 3.0;
  double c = 1.0 / std::sqrt(9.0 * d);
  for (;;) {
    double x = std_normal(rng);
    double v_cbrt = 1.0 + c * x;
    if (v_cbrt <= 0.0) continue;
    double v = v_cbrt * v_cbrt * v_cbrt;
    double u = open01_f64(rng);
    double x_sqr = x * x;
    if (u < 1.0 - 0.0331 * x_sqr * x_sqr || std::log(u) < 0.5 * x_sqr + d * (1.0 - v + std::log(v)))
      return d * v;
  }
}
static double sample_gamma(double alpha, SmallRng& rng) {
  if (alpha == 1.0) return -std::log(open01_f64(rng));
  if (alpha < 1.0) {
    double u = open01_f64(rng);
    return gamma_large(alpha + 1.0, rng) * std::pow(u, 1.0 / alpha);
  }
  return gamma_large(alpha, rng);
}
static void apply_dirichlet_noise(HalfNode& half, float epsilon, float concentration,
                                  SmallRng& rng) {
  int n = half.n_outcomes;
  if (n <= 1) return;
  double alpha = (double)(concentration / (float)n);
  if (!(alpha > 0.0)) return;
  float noise[5] = {0, 0, 0, 0, 0};
  float total = 0.f;
  for (int i = 0; i < n; ++i) {
    noise[i] = (float)sample_gamma(alpha, rng);
    total += noise[i];
  }
  if (total < std::numeric_limits<float>::min()) return;
  for (int i = 0; i < n; ++i)
    half.prior[i] = half.prior[i] * (1.0f - epsilon) + epsilon * noise[i] / total;
}

struct BatchStats {
  uint32_t nn_evals, terminals, collisions;
};

static int simulate_batch(MCTSTree& tree, const GameState& game, EvalFn eval, void* user,
                          const SearchConfig& cfg, uint32_t batch_size, SmallRng& rng,
                          BatchStats& out, SearchCounters* ctr) {  // search.rs:961-1073
  Node* root = tree.root;
  int32_t collisions_left = (int32_t)calculate_collisions_left(tree.node_count, cfg);
  std::vector<NodeToProcess> all_tp;
  all_tp.reserve(batch_size);
  std::vector<Collision> all_coll;
  uint32_t minibatch_size = 0, terminals = 0;

  while (minibatch_size < batch_size && collisions_left > 0) {
    uint32_t budget = std::min<uint32_t>((uint32_t)collisions_left, batch_size - minibatch_size);
    std::vector<NodeToProcess> tp;
    std::vector<Collision> coll;
    pick_nodes_to_extend(tree, game, cfg, budget, rng, tp, coll, ctr);
    for (auto& e : tp) {
      if (e.kind == Kind::Terminal) terminals += e.multivisit;
      minibatch_size += 1;
      all_tp.push_back(e);
    }
    for (auto& c : coll) collisions_left -= (int32_t)c.mv;
    all_coll.insert(all_coll.end(), coll.begin(), coll.end());
  }

  uint32_t nn_evals = 0, total_collisions = 0;
  std::vector<const GameState*> states;
  for (auto& e : all_tp)
    if (e.kind == Kind::NeedsEval) {
      nn_evals++;
      states.push_back(&e.state);
    }
  for (auto& c : all_coll) total_collisions += c.mv;

  std::vector<EvalResult> evals(states.size());
  if (!states.empty()) {
    int rc = eval(user, states.data(), (int)states.size(), evals.data());
    if (rc != 0) {  // GatherCleanupGuard, search.rs:945-955
      for (auto& e : all_tp) cancel_leaf_and_path(e.node, e.multivisit);
      cancel_shared_collisions(all_coll, root);
      return rc;
    }
  }

  size_t eval_idx = 0;
  for (auto& e : all_tp) {
    // roofline accounting: every backed-up entry touches depth+1 nodes
    if (ctr)
      for (Node* n = e.node; n; n = n->parent) ctr->path_nodes++;
    if (e.kind == Kind::NeedsEval) {
      const EvalResult& ev = evals[eval_idx++];
      populate_node(e.node, &ev);
      if (e.node == root && cfg.noise_epsilon > 0.0f) {
        apply_dirichlet_noise(e.node->p1, cfg.noise_epsilon, cfg.noise_concentration, rng);
        apply_dirichlet_noise(e.node->p2, cfg.noise_epsilon, cfg.noise_concentration, rng);
      }
      backup_and_finalize(e.node, ev.value_p1, ev.value_p2, e.multivisit, ctr);
    } else {
      backup_and_finalize(e.node, 0.f, 0.f, e.multivisit, ctr);
    }
  }
  cancel_shared_collisions(all_coll, root);
  out.nn_evals = nn_evals;
  out.terminals = terminals;
  out.collisions = total_collisions;
  return 0;
}

static void extract_half(const HalfNode& half, float node_value, float value_scale,
                         uint32_t children_visits, const SearchConfig& cfg, float policy[5],
                         float visit_counts[5], float& value) {  // search.rs:1116-1177
  int n = half.n_outcomes;
  for (int i = 0; i < 5; ++i) policy[i] = visit_counts[i] = 0.f;
  if (n == 0) {
    value = node_value;
    return;
  }
  float fpu = compute_fpu(half, node_value, value_scale, cfg.fpu_reduction);
  float q[5] = {0}, raw[5] = {0}, prior[5] = {0}, qn[5] = {0};
  for (int i = 0; i < n; ++i) {
    q[i] = half.edges[i].visits > 0 ? half.edges[i].q : fpu;
    raw[i] = (float)half.edges[i].visits;
    prior[i] = half.prior[i];
    qn[i] = q[i] / value_scale;
  }
  float pruned[5];
  compute_pruned_visits(qn, prior, raw, n, children_visits, cfg.c_puct, pruned);
  for (int i = 0; i < n; ++i) visit_counts[half.outcomes[i]] = pruned[i];
  float sum = 0.f;
  for (int i = 0; i < 5; ++i) {
    policy[i] = visit_counts[i];
    sum += policy[i];
  }
  if (sum > 0.0f) {
    for (int i = 0; i < 5; ++i) policy[i] /= sum;
  } else {
    half.expand_prior(policy);
  }
  float visit_sum = 0.f;
  for (int i = 0; i < n; ++i) visit_sum += raw[i];
  if (visit_sum > 0.0f) {
    float dot = 0.f;
    for (int i = 0; i < n; ++i) dot += q[i] * raw[i];
    value = dot / visit_sum;
  } else {
    value = node_value;
  }
}

int run_search(MCTSTree& tree, const GameState& game, EvalFn eval, void* user,
               const SearchConfig& cfg, uint32_t n_sims, uint32_t batch_size, SmallRng& rng,
               SearchResult& out, SearchCounters* ctr) {  // search.rs:362-390
  uint32_t remaining = n_sims, nn = 0, term = 0, coll = 0;
  while (remaining > 0) {
    BatchStats b{0, 0, 0};
    int rc = simulate_batch(tree, game, eval, user, cfg, std::min(remaining, batch_size), rng, b, ctr);
    if (rc != 0) return rc;
    nn += b.nn_evals;
    term += b.terminals;
    coll += b.collisions;
    uint32_t produced = std::max<uint32_t>(b.nn_evals + b.terminals, 1);
    remaining = remaining > produced ? remaining - produced : 0;
  }
  const Node* root = tree.root;  // extract_result, search.rs:1079-1111
  out.total_visits = root->total_visits;
  uint32_t cv = root->children_visits();
  extract_half(root->p1, root->v1, root->value_scale, cv, cfg, out.policy_p1, out.visit_counts_p1,
               out.value_p1);
  extract_half(root->p2, root->v2, root->value_scale, cv, cfg, out.policy_p2, out.visit_counts_p2,
               out.value_p2);
  root->p1.expand_prior(out.prior_p1);
  root->p2.expand_prior(out.prior_p2);
  out.nn_evals = nn;
  out.terminals = term;
  out.collisions = coll;
  return 0;
}

bool tree_all_in_flight_zero(const Node* root) {  // invariant of search.rs:2750-2791
  std::vector<const Node*> stack{root};
  while (!stack.empty()) {
    const Node* n = stack.back();
    stack.pop_back();
    if (n->n_in_flight != 0) return false;
    for (int i = 0; i < n->p1.n_outcomes; ++i)
      if (n->p1.edges[i].n_in_flight != 0) return false;
    for (int i = 0; i < n->p2.n_outcomes; ++i)
      if (n->p2.edges[i].n_in_flight != 0) return false;
    if (n->first_child) stack.push_back(n->first_child);
    if (n->next_sibling) stack.push_back(n->next_sibling);
  }
  return true;
}

// =======================================================================================
// selfplay.rs
// =======================================================================================
static void cheese_bits(const GameState& g, uint8_t out[AR_MAX_CELLS / 8]) {
  std::memset(out, 0, AR_MAX_CELLS / 8);
  int cells = g.width * g.height;
  for (int c = 0; c < cells; ++c)
    if (g.has_cheese(c)) out[c >> 3] |= (uint8_t)(1u << (c & 7));
}

static void fill_search_result(ar_search_result& o, const SearchResult& r, const MCTSTree& tree) {
  std::memset(&o, 0, sizeof(o));
  std::memcpy(o.policy_p1, r.policy_p1, 20);
  std::memcpy(o.policy_p2, r.policy_p2, 20);
  o.value_p1 = r.value_p1;
  o.value_p2 = r.value_p2;
  std::memcpy(o.visit_counts_p1, r.visit_counts_p1, 20);
  std::memcpy(o.visit_counts_p2, r.visit_counts_p2, 20);
  std::memcpy(o.prior_p1, r.prior_p1, 20);
  std::memcpy(o.prior_p2, r.prior_p2, 20);
  o.total_visits = r.total_visits;
  o.nn_evals = r.nn_evals;
  o.terminals = r.terminals;
  o.collisions = r.collisions;
  float ev[5];
  tree.root->p1.expand_visits(ev);
  for (int i = 0; i < 5; ++i) o.raw_visits_p1[i] = (uint32_t)ev[i];
  tree.root->p2.expand_visits(ev);
  for (int i = 0; i < 5; ++i) o.raw_visits_p2[i] = (uint32_t)ev[i];
  o.node_count = tree.node_count;
}

void fill_search_result_public(ar_search_result& o, const SearchResult& r, const MCTSTree& t) {
  fill_search_result(o, r, t);
}

int play_game(GameState game, EvalFn eval, void* user, const SearchConfig& cfg, uint32_t n_sims,
              uint32_t batch_size, SmallRng& rng, uint32_t game_index, NodeArena* arena,
              ar_game_summary& summary, ar_position_record* positions, int positions_cap,
              SearchCounters* ctr) {  // selfplay.rs:515-598
  std::memset(&summary, 0, sizeof(summary));
  summary.game_index = game_index;
  summary.cheese_available = game.remaining_cheese();
  MCTSTree tree(game, arena);
  int n_pos = 0;
  while (!game.check_game_over()) {
    SearchResult r;
    int rc = run_search(tree, game, eval, user, cfg, n_sims, batch_size, rng, r, ctr);
    if (rc != 0) return rc;
    summary.total_simulations += r.total_visits;
    summary.total_nn_evals += r.nn_evals;
    summary.total_terminals += r.terminals;
    summary.total_collisions += r.collisions;
    uint8_t a1 = sample_action(r.policy_p1, rng);
    uint8_t a2 = sample_action(r.policy_p2, rng);
    if (n_pos >= positions_cap) return -2;
    ar_position_record& p = positions[n_pos++];  // record_position, selfplay.rs:486-512
    std::memset(&p, 0, sizeof(p));
    p.p1_x = game.p1x; p.p1_y = game.p1y; p.p2_x = game.p2x; p.p2_y = game.p2y;
    p.p1_mud = game.mud1; p.p2_mud = game.mud2;
    p.action_p1 = a1; p.action_p2 = a2;
    p.turn = game.turn;
    p.p1_score = game.s1; p.p2_score = game.s2;
    fill_search_result(p.search, r, tree);
    cheese_bits(game, p.cheese);
    game.make_move(a1, a2);
    if (!tree.advance_root(a1, a2)) tree.reinit(game);
  }
  summary.n_positions = (uint32_t)n_pos;
  summary.final_p1_score = game.s1;
  summary.final_p2_score = game.s2;
  summary.result = game.s1 > game.s2 ? 1 : (game.s2 > game.s1 ? 2 : 0);
  // compute_cheese_outcomes, selfplay.rs:415-471
  int cells = game.width * game.height;
  for (int c = 0; c < AR_MAX_CELLS; ++c) summary.cheese_outcomes[c] = 2;
  uint8_t final_bits[AR_MAX_CELLS / 8];
  cheese_bits(game, final_bits);
  for (int i = 0; i < n_pos; ++i) {
    const uint8_t* cur = positions[i].cheese;
    const uint8_t* nxt = (i + 1 < n_pos) ? positions[i + 1].cheese : final_bits;
    int n1x, n1y, n2x, n2y;
    if (i + 1 < n_pos) {
      n1x = positions[i + 1].p1_x; n1y = positions[i + 1].p1_y;
      n2x = positions[i + 1].p2_x; n2y = positions[i + 1].p2_y;
    } else {
      n1x = game.p1x; n1y = game.p1y; n2x = game.p2x; n2y = game.p2y;
    }
    for (int c = 0; c < cells; ++c) {
      bool a = (cur[c >> 3] >> (c & 7)) & 1, b = (nxt[c >> 3] >> (c & 7)) & 1;
      if (a && !b) {
        int x = c % game.width, y = c / game.width;
        bool p1t = (n1x == x && n1y == y), p2t = (n2x == x && n2y == y);
        summary.cheese_outcomes[c] = (p1t && p2t) ? 1 : p1t ? 0 : p2t ? 3 : 2;
      }
    }
  }
  return 0;
}

// =======================================================================================
// flat_encoder.rs:52-124
// =======================================================================================
int obs_dim(int w, int h) { return w * h * 7 + 6; }

void encode_flat(const GameState& g, float* out) {
  const float MAX_MUD_COST = 10.0f, MAX_MUD_TURNS = 10.0f, MAX_SCORE = 10.0f;
  int w = g.width, h = g.height, spatial = w * h;
  for (int i = 0; i < spatial * 4; ++i) out[i] = -1.0f;
  for (int c = 0; c < spatial; ++c)
    for (int d = 0; d < 4; ++d) {
      uint8_t cost = g.cost(c, d);
      if (cost == 0) continue;
      out[c * 4 + d] = cost >= 2 ? (float)cost / MAX_MUD_COST : 1.0f / MAX_MUD_COST;
    }
  for (int j = 0; j < spatial * 3; ++j) out[spatial * 4 + j] = 0.0f;
  out[spatial * 4 + g.cell(g.p1x, g.p1y)] = 1.0f;
  out[spatial * 5 + g.cell(g.p2x, g.p2y)] = 1.0f;
  for (int c = 0; c < spatial; ++c) out[spatial * 6 + c] = g.has_cheese(c) ? 1.0f : 0.0f;
  float* s = out + spatial * 7;
  s[0] = g.s1 - g.s2;
  s[1] = g.max_turns > 0 ? (float)g.turn / (float)g.max_turns : 0.0f;
  s[2] = (float)g.mud1 / MAX_MUD_TURNS;
  s[3] = (float)g.mud2 / MAX_MUD_TURNS;
  s[4] = g.s1 / MAX_SCORE;
  s[5] = g.s2 / MAX_SCORE;
}

}  // namespace orc
