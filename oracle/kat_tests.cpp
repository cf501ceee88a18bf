// ORACLE — TEST INFRASTRUCTURE ONLY.
// Known-answer tests that pin the oracle against the values the reference's own unit tests
// assert (file:line cited per test).  Run by tests/test_oracle_kat.py; exit code 0 = all pass.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "alpharat_oracle.hpp"

using namespace orc;

static int g_fail = 0, g_pass = 0;
#define CHECK(cond)                                                         \
  do {                                                                      \
    if (!(cond)) {                                                          \
      std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond);           \
      g_fail++;                                                             \
    } else {                                                                \
      g_pass++;                                                             \
    }                                                                       \
  } while (0)
#define NEAR(a, b, tol) CHECK(std::fabs((double)(a) - (double)(b)) < (tol))

static Node* open_node(NodeArena* a) {  // search.rs:1198-1208 make_open_node
  static const float pr[5] = {0.2f, 0.2f, 0.2f, 0.2f, 0.2f};
  static const uint8_t eff[5] = {0, 1, 2, 3, 4};
  Node* n = new Node();
  (void)a;
  n->p1 = HalfNode::make(pr, eff);
  n->p2 = HalfNode::make(pr, eff);
  return n;
}
static void wire(Node* parent, Node* child, uint8_t i, uint8_t j, float r1, float r2, float scale) {
  child->parent = parent;
  child->po1 = i;
  child->po2 = j;
  child->edge_r1 = r1;
  child->edge_r2 = r2;
  child->value_scale = scale;
  child->next_sibling = parent->first_child;
  parent->first_child = child;
}

static GameState open_game(int w, int h, int p1x, int p1y, int p2x, int p2y,
                           std::vector<std::pair<int, int>> cheese, int max_turns) {
  GameState g;
  g.width = (uint8_t)w; g.height = (uint8_t)h; g.max_turns = (uint16_t)max_turns;
  g.p1x = (uint8_t)p1x; g.p1y = (uint8_t)p1y; g.p2x = (uint8_t)p2x; g.p2y = (uint8_t)p2y;
  g.maze = std::make_shared<MazeData>();
  std::memset(g.maze->move_cost, 0, sizeof(g.maze->move_cost));
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      int c = y * w + x;
      g.maze->move_cost[c * 4 + 0] = y + 1 < h;
      g.maze->move_cost[c * 4 + 1] = x + 1 < w;
      g.maze->move_cost[c * 4 + 2] = y > 0;
      g.maze->move_cost[c * 4 + 3] = x > 0;
    }
  for (auto& c : cheese) {
    g.set_cheese(c.second * w + c.first, true);
    g.remaining++;
  }
  return g;
}
static void add_wall(GameState& g, int x1, int y1, int x2, int y2, uint8_t val = 0) {
  int d = (x2 == x1) ? (y2 > y1 ? 0 : 2) : (x2 > x1 ? 1 : 3);
  g.maze->move_cost[(y1 * g.width + x1) * 4 + d] = val;
  g.maze->move_cost[(y2 * g.width + x2) * 4 + ((d + 2) % 4)] = val;
}

int main() {
  NodeArena* arena = arena_new();

  // ---- rand: xoshiro256++ reference vector (xoshiro256plusplus.c / rand_xoshiro tests) ----
  {
    SmallRng r = SmallRng::from_state(1, 2, 3, 4);
    const uint64_t exp[10] = {41943041ULL, 58720359ULL, 3588806011781223ULL, 3591011842654386ULL,
                              9228616714210784205ULL, 9973669472204895162ULL,
                              14011001112246962877ULL, 12406186145184390807ULL,
                              15849039046786891736ULL, 10450023813501588000ULL};
    for (int i = 0; i < 10; ++i) CHECK(r.next_u64() == exp[i]);
  }
  {  // SplitMix64 reference vector for seed 0 (first output 0xe220a8397b1dcdaf)
    SmallRng r = SmallRng::seed_from_u64(0);
    CHECK(r.s[0] == 0xe220a8397b1dcdafULL);
    CHECK(r.s[1] == 0x6e789e6aa1b965f4ULL);
    CHECK(r.s[2] == 0x06c45d188009454fULL);
    CHECK(r.s[3] == 0xf88bb8a8724c81ecULL);
  }
  {  // gen_range stays in range; n=1 consumes one draw and returns 0
    SmallRng r = SmallRng::seed_from_u64(42);
    for (uint32_t n = 1; n < 40; ++n)
      for (int k = 0; k < 50; ++k) CHECK(r.gen_range_u32(n) < n);
    for (int k = 0; k < 200; ++k) {
      float x = r.uniform_f32(0.0f, 1.0f);
      CHECK(x >= 0.0f && x < 1.0f);
    }
  }
  {  // sample_action degenerate cases — selfplay.rs:860-905
    SmallRng r = SmallRng::seed_from_u64(7);
    const float zero[5] = {0, 0, 0, 0, 0};
    CHECK(sample_action(zero, r) == 4);
    for (int a = 0; a < 5; ++a) {
      float onehot[5] = {0, 0, 0, 0, 0};
      onehot[a] = 1.0f;
      for (int k = 0; k < 20; ++k) CHECK(sample_action(onehot, r) == a);
    }
    const float two[5] = {0.5f, 0, 0.5f, 0, 0};
    int c0 = 0, c2 = 0;
    for (int k = 0; k < 2000; ++k) {
      uint8_t a = sample_action(two, r);
      CHECK(a == 0 || a == 2);
      c0 += a == 0; c2 += a == 2;
    }
    CHECK(c0 > 800 && c2 > 800);
  }

  // ---- compute_outcomes tables — node.rs:496-543 ----
  {
    uint8_t out[5], a2i[5], n;
    const uint8_t open[5] = {0, 1, 2, 3, 4};
    compute_outcomes(open, out, n, a2i);
    CHECK(n == 5);
    for (int a = 0; a < 5; ++a) CHECK(out[a] == a && a2i[a] == a);
    const uint8_t wall[5] = {4, 1, 2, 3, 4};
    compute_outcomes(wall, out, n, a2i);
    CHECK(n == 4 && out[0] == 1 && out[1] == 2 && out[2] == 3 && out[3] == 4);
    CHECK(a2i[0] == a2i[4] && out[a2i[0]] == 4);
    const uint8_t corner[5] = {4, 1, 2, 4, 4};
    compute_outcomes(corner, out, n, a2i);
    CHECK(n == 3 && out[0] == 1 && out[1] == 2 && out[2] == 4);
    CHECK(a2i[0] == a2i[3] && a2i[0] == a2i[4]);
    const uint8_t mud[5] = {4, 4, 4, 4, 4};
    compute_outcomes(mud, out, n, a2i);
    CHECK(n == 1 && out[0] == 4);
    for (int a = 0; a < 5; ++a) CHECK(a2i[a] == 0);
  }
  // ---- prior reduction — node.rs:548-586 ----
  {
    const float u[5] = {0.2f, 0.2f, 0.2f, 0.2f, 0.2f};
    const uint8_t wall[5] = {4, 1, 2, 3, 4};
    HalfNode h = HalfNode::make(u, wall);
    CHECK(h.n_outcomes == 4);
    NEAR(h.prior[h.action_to_idx[4]], 0.4, 1e-6);
    const float nu[5] = {0.1f, 0.3f, 0.2f, 0.15f, 0.25f};
    h = HalfNode::make(nu, wall);
    NEAR(h.prior[h.action_to_idx[4]], 0.35, 1e-6);
    const uint8_t mud[5] = {4, 4, 4, 4, 4};
    h = HalfNode::make(nu, mud);
    CHECK(h.n_outcomes == 1);
    NEAR(h.prior[0], 1.0, 1e-6);
    // expand_visits — node.rs:590-612
    h = HalfNode::make(u, wall);
    h.edges[0].visits = 10;
    h.edges[3].visits = 7;
    float ev[5];
    h.expand_visits(ev);
    CHECK(ev[0] == 0 && ev[1] == 10 && ev[2] == 0 && ev[3] == 0 && ev[4] == 7);
  }
  // ---- smart_uniform_prior — tree.rs:432-461 ----
  {
    float p[5];
    const uint8_t open[5] = {0, 1, 2, 3, 4};
    smart_uniform_prior(open, p);
    for (int a = 0; a < 5; ++a) NEAR(p[a], 0.2, 1e-6);
    const uint8_t wall[5] = {4, 1, 2, 3, 4};
    smart_uniform_prior(wall, p);
    CHECK(p[0] == 0.0f);
    for (int a = 1; a < 5; ++a) NEAR(p[a], 0.25, 1e-6);
    const uint8_t mud[5] = {4, 4, 4, 4, 4};
    smart_uniform_prior(mud, p);
    CHECK(p[4] == 1.0f && p[0] == 0.0f && p[1] == 0.0f && p[2] == 0.0f && p[3] == 0.0f);
  }
  // ---- virtual-loss state machine — node.rs:914-982 ----
  {
    Node n;
    CHECK(n.try_start_score_update());   // unvisited, first claim succeeds
    CHECK(!n.try_start_score_update());  // unvisited and in flight: collision
    CHECK(n.n_in_flight == 1);
    n.finalize_score_update(1.0f, 1.0f, 1);
    CHECK(n.n_in_flight == 0 && n.total_visits == 1);
    CHECK(n.try_start_score_update());  // visited: always succeeds
    CHECK(n.try_start_score_update());
    CHECK(n.n_in_flight == 2);
  }
  // ---- multivisit == repeated single — node.rs:986-1071, search.rs:1945-2000 ----
  {
    Node m, s;
    m.n_in_flight = 5;
    m.finalize_score_update(3.0f, 5.0f, 5);
    for (int i = 0; i < 5; ++i) s.update_value(3.0f, 5.0f);
    CHECK(m.total_visits == s.total_visits);
    NEAR(m.v1, s.v1, 1e-6);
    NEAR(m.v2, s.v2, 1e-6);
    CHECK(m.n_in_flight == 0);
    HalfEdge em, es;
    em.update_multivisit(7.0f, 4);
    for (int i = 0; i < 4; ++i) es.update(7.0f);
    CHECK(em.visits == es.visits);
    NEAR(em.q, es.q, 1e-6);
    Node m2, s2;
    m2.n_in_flight = 5;
    m2.finalize_score_update(2.0f, 2.0f, 3);
    m2.finalize_score_update(8.0f, 8.0f, 2);
    for (int i = 0; i < 3; ++i) s2.update_value(2.0f, 2.0f);
    for (int i = 0; i < 2; ++i) s2.update_value(8.0f, 8.0f);
    NEAR(m2.v1, s2.v1, 1e-5);
    NEAR(m2.v1, 4.4, 1e-5);
  }
  // ---- backup KATs — search.rs:1216-1489, 1880-1944 ----
  {  // backup_single_level
    Node* root = open_node(arena);
    Node* child = open_node(arena);
    root->value_scale = 5.0f;
    wire(root, child, 0, 1, 1.0f, 0.5f, 5.0f);
    backup({root}, {0}, {1}, child, 3.0f, 2.0f);
    CHECK(child->total_visits == 1);
    NEAR(child->v1, 3.0, 1e-6); NEAR(child->v2, 2.0, 1e-6);
    CHECK(root->total_visits == 1);
    NEAR(root->v1, 4.0, 1e-6); NEAR(root->v2, 2.5, 1e-6);
    CHECK(root->p1.edges[0].visits == 1); NEAR(root->p1.edges[0].q, 4.0, 1e-6);
    CHECK(root->p2.edges[1].visits == 1); NEAR(root->p2.edges[1].q, 2.5, 1e-6);
  }
  {  // backup_two_level_q_chain
    Node *root = open_node(arena), *mid = open_node(arena), *leaf = open_node(arena);
    root->value_scale = 5.0f;
    wire(root, mid, 0, 0, 1.0f, 0.5f, 5.0f);
    wire(mid, leaf, 1, 2, 0.5f, 1.0f, 5.0f);
    backup({root, mid}, {0, 1}, {0, 2}, leaf, 2.0f, 3.0f);
    NEAR(leaf->v1, 2.0, 1e-6); NEAR(leaf->v2, 3.0, 1e-6);
    NEAR(mid->v1, 2.5, 1e-6); NEAR(mid->v2, 4.0, 1e-6);
    NEAR(root->v1, 3.5, 1e-6); NEAR(root->v2, 4.5, 1e-6);
  }
  {  // backup_multiple_same_edge + backup_value_mixing
    Node *root = open_node(arena), *child = open_node(arena);
    root->value_scale = 5.0f;
    wire(root, child, 0, 0, 0.0f, 0.0f, 5.0f);
    backup({root}, {0}, {0}, child, 2.0f, 1.0f);
    backup({root}, {0}, {0}, child, 4.0f, 3.0f);
    backup({root}, {0}, {0}, child, 6.0f, 5.0f);
    CHECK(child->total_visits == 3);
    NEAR(child->v1, 4.0, 1e-5); NEAR(child->v2, 3.0, 1e-5);
    CHECK(root->p1.edges[0].visits == 3);
    NEAR(root->p1.edges[0].q, 4.0, 1e-5); NEAR(root->p2.edges[0].q, 3.0, 1e-5);
    Node *r2 = open_node(arena), *c2 = open_node(arena);
    r2->value_scale = 5.0f;
    wire(r2, c2, 0, 0, 1.0f, 0.0f, 5.0f);
    backup({r2}, {0}, {0}, c2, 2.0f, 0.0f);
    backup({r2}, {0}, {0}, c2, 4.0f, 0.0f);
    backup({r2}, {0}, {0}, c2, 6.0f, 0.0f);
    NEAR(r2->v1, 5.0, 1e-5);
  }
  {  // backup_multi_level_multi_backup (the "killer regression", search.rs:1945-2000)
    Node *root = open_node(arena), *mid = open_node(arena), *leaf = open_node(arena);
    root->value_scale = 15.0f;
    wire(root, mid, 0, 0, 1.0f, 1.0f, 15.0f);
    wire(mid, leaf, 0, 0, 0.5f, 0.5f, 15.0f);
    backup({root, mid}, {0, 0}, {0, 0}, leaf, 10.0f, 10.0f);
    NEAR(leaf->v1, 10.0, 1e-5); NEAR(mid->v1, 10.5, 1e-5); NEAR(root->v1, 11.5, 1e-5);
    backup({root, mid}, {0, 0}, {0, 0}, leaf, 6.0f, 6.0f);
    NEAR(leaf->v1, 8.0, 1e-5); NEAR(mid->v1, 8.5, 1e-5); NEAR(root->v1, 9.5, 1e-5);
  }
  {  // backup_and_finalize == backup for multivisit 1, and clears virtual losses
    Node *root = open_node(arena), *mid = open_node(arena), *leaf = open_node(arena);
    root->value_scale = 15.0f;
    wire(root, mid, 2, 3, 1.0f, 0.0f, 15.0f);
    wire(mid, leaf, 1, 4, 0.5f, 0.5f, 15.0f);
    root->n_in_flight = 1; mid->n_in_flight = 1; leaf->n_in_flight = 1;
    root->p1.edges[2].n_in_flight = 1; root->p2.edges[3].n_in_flight = 1;
    mid->p1.edges[1].n_in_flight = 1; mid->p2.edges[4].n_in_flight = 1;
    backup_and_finalize(leaf, 2.0f, 3.0f, 1, nullptr);
    NEAR(leaf->v1, 2.0, 1e-6); NEAR(mid->v1, 2.5, 1e-6); NEAR(root->v1, 3.5, 1e-6);
    NEAR(mid->v2, 3.5, 1e-6); NEAR(root->v2, 3.5, 1e-6);
    CHECK(tree_all_in_flight_zero(root));
    CHECK(root->p1.edges[2].visits == 1 && root->p2.edges[3].visits == 1);
  }
  // ---- compute_pruned_visits — search.rs:1773-1875 ----
  {
    float r[5];
    const float pr[5] = {0.2f, 0.2f, 0.2f, 0.2f, 0.2f};
    const float q1[5] = {0.5f, 0.3f, 0.8f, 0.2f, 0.1f};
    const float v1[5] = {10, 5, 20, 3, 2};
    compute_pruned_visits(q1, pr, v1, 5, 40, 1.5f, r);
    NEAR(r[2], 20.0, 1e-6);
    for (int i = 0; i < 5; ++i) CHECK(r[i] <= v1[i] + 1e-6f && r[i] >= 0.0f);
    const float q2[5] = {0.5f, 0.3f, 0.8f, 0.95f, 0.1f};
    const float v2[5] = {10, 5, 20, 18, 2};
    compute_pruned_visits(q2, pr, v2, 5, 55, 1.5f, r);
    NEAR(r[3], 18.0, 1e-6);
    const float q3[5] = {0.8f, 0.1f, 0, 0, 0};
    const float v3[5] = {50, 20, 10, 10, 10};
    compute_pruned_visits(q3, pr, v3, 5, 100, 1.5f, r);
    NEAR(r[0], 50.0, 1e-6);
    for (int i = 2; i < 5; ++i) CHECK(r[i] <= v3[i]);
    // exact value of a capped entry: c*p*sqrt(100)/(puct*-0) - 1, puct* = 0.8 + 1.5*0.2*10/51
    float puct_star = 0.8f + 1.5f * 0.2f * 10.0f / 51.0f;
    NEAR(r[2], 1.5f * 0.2f * 10.0f / puct_star - 1.0f, 1e-5);
    const float q4[5] = {0.9f, 0, 0, 0, 0};
    const float v4[5] = {50, 1, 1, 1, 1};
    compute_pruned_visits(q4, pr, v4, 5, 54, 1.5f, r);
    for (int i = 0; i < 5; ++i) CHECK(r[i] >= 0.0f);
    const float q5[1] = {0.5f}, p5[1] = {1.0f}, v5[1] = {42.0f};
    compute_pruned_visits(q5, p5, v5, 1, 42, 1.5f, r);
    NEAR(r[0], 42.0, 1e-6);
  }
  // ---- calculate_collisions_left — search.rs:3578-3655 ----
  {
    SearchConfig c;
    c.collision_limit_min = 2; c.collision_limit_max = 128;
    CHECK(calculate_collisions_left(0, c) == 2);
    CHECK(calculate_collisions_left(799, c) == 2);
    CHECK(calculate_collisions_left(800, c) == 2);
    CHECK(calculate_collisions_left(50000, c) == 128);
    CHECK(calculate_collisions_left(100000, c) == 128);
    SearchConfig l;
    l.collision_limit_min = 0; l.collision_limit_max = 100;
    l.collision_scaling_start = 0; l.collision_scaling_end = 100;
    CHECK(calculate_collisions_left(50, l) == 50);
    CHECK(calculate_collisions_left(25, l) == 25);
    l.collision_scaling_power = 2.0f;
    CHECK(calculate_collisions_left(50, l) == 25);
    CHECK(calculate_collisions_left(100, l) == 100);
    SearchConfig e;
    e.collision_limit_min = 5; e.collision_limit_max = 200;
    e.collision_scaling_start = 1000; e.collision_scaling_end = 1000;
    CHECK(calculate_collisions_left(999, e) == 5);
    CHECK(calculate_collisions_left(1000, e) == 200);
    CHECK(calculate_collisions_left(1001, e) == 200);
  }
  // ---- game-step rewards through the engine — tree.rs:934-999 ----
  {
    GameState g = open_game(5, 5, 0, 0, 4, 4, {{1, 0}}, 100);  // one_cheese_adjacent_game
    g.make_move(1, 4);
    CHECK(g.s1 == 1.0f && g.s2 == 0.0f && g.remaining == 0);
    g = open_game(5, 5, 0, 0, 4, 4, {{1, 0}}, 100);
    g.make_move(0, 4);
    CHECK(g.s1 == 0.0f && g.s2 == 0.0f && g.p1x == 0 && g.p1y == 1);
    g = open_game(5, 5, 0, 0, 2, 0, {{1, 0}}, 100);  // contested_cheese_game
    g.make_move(1, 3);
    CHECK(g.s1 == 0.5f && g.s2 == 0.5f && g.remaining == 0);
    g = open_game(5, 5, 0, 0, 4, 0, {{1, 0}, {3, 0}}, 100);
    g.make_move(1, 3);
    CHECK(g.s1 == 1.0f && g.s2 == 1.0f);
    g = open_game(5, 5, 0, 0, 2, 0, {{1, 0}}, 100);
    g.make_move(4, 3);
    CHECK(g.s1 == 0.0f && g.s2 == 1.0f);
  }
  // ---- effective actions — backend.rs:148-256, tree.rs:465-547, search.rs:2487-2508 ----
  {
    uint8_t e[5];
    GameState g = open_game(5, 5, 2, 2, 0, 0, {{4, 4}}, 100);
    g.effective_actions_p1(e);  // centre: identity
    for (int a = 0; a < 5; ++a) CHECK(e[a] == a);
    g.effective_actions_p2(e);  // bottom-left corner: DOWN, LEFT blocked
    CHECK(e[0] == 0 && e[1] == 1 && e[2] == 4 && e[3] == 4 && e[4] == 4);
    g = open_game(5, 5, 4, 4, 0, 2, {{2, 2}}, 100);
    g.effective_actions_p1(e);  // top-right corner: UP, RIGHT blocked
    CHECK(e[0] == 4 && e[1] == 4 && e[2] == 2 && e[3] == 3 && e[4] == 4);
    g.effective_actions_p2(e);  // left edge
    CHECK(e[0] == 0 && e[1] == 1 && e[2] == 2 && e[3] == 4);
    g = open_game(5, 5, 2, 2, 0, 0, {{4, 4}}, 100);
    add_wall(g, 2, 2, 2, 3);
    g.effective_actions_p1(e);
    CHECK(e[0] == 4 && e[1] == 1 && e[2] == 2 && e[3] == 3);
    // corridor_game (test_util.rs:79-103): row 0 walled from row 1
    g = open_game(5, 5, 0, 0, 4, 0, {{2, 0}}, 100);
    for (int x = 0; x < 5; ++x) add_wall(g, x, 0, x, 1);
    g.effective_actions_p1(e);
    CHECK(e[0] == 4 && e[1] == 1 && e[2] == 4 && e[3] == 4);
    g.effective_actions_p2(e);
    CHECK(e[0] == 4 && e[1] == 4 && e[2] == 4 && e[3] == 3);
    // mud_game_p1_stuck (test_util.rs:32-49): all actions -> STAY
    g = open_game(5, 5, 2, 2, 4, 4, {{0, 0}}, 100);
    add_wall(g, 2, 2, 2, 3, 3);
    g.make_move(0, 4);
    CHECK(g.mud1 > 0);
    CHECK(g.p1x == 2 && g.p1y == 3);  // fixtures/mud_stuck_5x5.json: position = target
    g.effective_actions_p1(e);
    for (int a = 0; a < 5; ++a) CHECK(e[a] == 4);
  }
  // ---- termination — test_util.rs:105-117, eval/game.py:31-44 ----
  {
    GameState g = open_game(5, 5, 0, 0, 0, 1, {{4, 4}}, 1);
    CHECK(!g.check_game_over());
    g.make_move(4, 4);
    CHECK(g.check_game_over());
    g = open_game(5, 5, 0, 0, 4, 4, {{1, 0}, {2, 0}, {3, 3}}, 100);
    g.make_move(1, 4);
    CHECK(!g.check_game_over());  // 1 of 3
    g.make_move(1, 4);
    CHECK(g.check_game_over());  // 2 of 3 > 1.5: majority
  }
  // ---- end-to-end search invariants — search.rs:2437-2474,2570-2640,3137-3156,3690-3771 ----
  {
    SearchConfig cfg;
    for (uint32_t sims : {10u, 50u, 100u, 200u}) {  // python/tests/test_search.py:131-135
      GameState g = open_game(5, 5, 2, 2, 2, 2, {{0, 0}, {4, 4}, {0, 4}, {4, 0}, {1, 3}}, 100);
      SmallRng rng = SmallRng::seed_from_u64(42);
      MCTSTree tree(g, arena);
      SearchResult r;
      CHECK(run_search(tree, g, smart_uniform_eval, nullptr, cfg, sims, 8, rng, r, nullptr) == 0);
      CHECK(r.total_visits == sims);
      CHECK(r.total_visits == r.nn_evals + r.terminals);
      uint32_t e1 = 0, e2 = 0;
      for (int i = 0; i < 5; ++i) { e1 += tree.root->p1.edges[i].visits; e2 += tree.root->p2.edges[i].visits; }
      CHECK(e1 == r.total_visits - 1 && e2 == r.total_visits - 1);
      CHECK(tree_all_in_flight_zero(tree.root));
      float s1 = 0, s2 = 0;
      for (int i = 0; i < 5; ++i) { s1 += r.policy_p1[i]; s2 += r.policy_p2[i]; }
      NEAR(s1, 1.0, 1e-5); NEAR(s2, 1.0, 1e-5);
    }
    {  // batch_size 1: total_visits == n_sims (search.rs:2437-2451)
      GameState g = open_game(5, 5, 0, 0, 4, 4, {{2, 2}, {1, 1}, {3, 3}}, 100);
      SmallRng rng = SmallRng::seed_from_u64(1);
      MCTSTree tree(g, arena);
      SearchResult r;
      run_search(tree, g, smart_uniform_eval, nullptr, cfg, 100, 1, rng, r, nullptr);
      CHECK(r.total_visits == 100);
      // blocked actions get exactly zero (search.rs:2624-2640): P1 in bottom-left corner
      CHECK(r.policy_p1[2] == 0.0f && r.policy_p1[3] == 0.0f);
      CHECK(r.visit_counts_p1[2] == 0.0f && r.visit_counts_p1[3] == 0.0f);
    }
    {  // stuck in mud => 100 % STAY (search.rs:2570-2588)
      GameState g = open_game(5, 5, 2, 2, 4, 4, {{0, 0}}, 100);
      add_wall(g, 2, 2, 2, 3, 3);
      g.make_move(0, 4);
      SmallRng rng = SmallRng::seed_from_u64(3);
      MCTSTree tree(g, arena);
      SearchResult r;
      run_search(tree, g, smart_uniform_eval, nullptr, cfg, 50, 8, rng, r, nullptr);
      NEAR(r.policy_p1[4], 1.0, 1e-6);
    }
    {  // terminal root: visits=100, terminals=100, nn_evals=0 (search.rs:3690-3718)
      GameState g = open_game(5, 5, 0, 0, 0, 1, {{4, 4}}, 1);
      g.make_move(4, 4);
      SmallRng rng = SmallRng::seed_from_u64(5);
      MCTSTree tree(g, arena);
      SearchResult r;
      run_search(tree, g, smart_uniform_eval, nullptr, cfg, 100, 8, rng, r, nullptr);
      CHECK(r.total_visits == 100 && r.terminals == 100 && r.nn_evals == 0);
      CHECK(tree_all_in_flight_zero(tree.root));
    }
    {  // visit accounting with tree reuse: post = pre + nn + term (search.rs:3741-3771)
      GameState g = open_game(5, 5, 0, 0, 2, 0, {{1, 0}}, 3);  // short_game
      SmallRng rng = SmallRng::seed_from_u64(9);
      MCTSTree tree(g, arena);
      SearchResult r;
      run_search(tree, g, smart_uniform_eval, nullptr, cfg, 64, 8, rng, r, nullptr);
      uint32_t pre = tree.root->total_visits;
      run_search(tree, g, smart_uniform_eval, nullptr, cfg, 64, 8, rng, r, nullptr);
      CHECK(r.total_visits == pre + r.nn_evals + r.terminals);
      CHECK(r.terminals > 0);
      CHECK(tree_all_in_flight_zero(tree.root));
    }
    {  // same seed => identical result (python/tests/test_search.py:111-119)
      GameState g = open_game(5, 5, 0, 0, 4, 4, {{2, 2}, {1, 3}, {3, 1}}, 100);
      SearchResult a, b;
      for (int rep = 0; rep < 2; ++rep) {
        SmallRng rng = SmallRng::seed_from_u64(1234);
        MCTSTree tree(g, arena);
        run_search(tree, g, smart_uniform_eval, nullptr, cfg, 200, 8, rng, rep ? b : a, nullptr);
      }
      CHECK(std::memcmp(&a, &b, sizeof(a)) == 0);
    }
    {  // advance_root with a blocked action reuses the STAY child (tree.rs:573-652)
      GameState g = open_game(5, 5, 0, 0, 4, 4, {{2, 2}}, 100);
      SmallRng rng = SmallRng::seed_from_u64(11);
      MCTSTree tree(g, arena);
      SearchResult r;
      run_search(tree, g, smart_uniform_eval, nullptr, cfg, 200, 8, rng, r, nullptr);
      uint8_t stay1 = tree.root->p1.action_to_idx[4], stay2 = tree.root->p2.action_to_idx[4];
      Node* stay_child = find_child(tree.root, stay1, stay2);
      CHECK(stay_child != nullptr);
      uint32_t expect_visits = stay_child ? stay_child->total_visits : 0;
      CHECK(tree.advance_root(2, 0));  // P1 DOWN (blocked) / P2 UP (blocked) == STAY/STAY
      CHECK(tree.root == stay_child && tree.root->total_visits == expect_visits);
      CHECK(tree.root->parent == nullptr);
    }
  }
  // ---- cheese-outcome attribution through play_game — selfplay.rs:1254-1413 ----
  {
    SearchConfig cfg;
    GameState g = open_game(3, 3, 0, 0, 2, 2, {{1, 0}, {1, 2}, {1, 1}}, 20);
    SmallRng rng = SmallRng::seed_from_u64(2);
    ar_game_summary s;
    std::vector<ar_position_record> pos(20);
    CHECK(play_game(g, smart_uniform_eval, nullptr, cfg, 50, 8, rng, 7, arena, s, pos.data(), 20, nullptr) == 0);
    CHECK(s.game_index == 7 && s.n_positions >= 1 && s.n_positions <= 20);
    CHECK(s.cheese_available == 3);
    float collected = 0;
    for (int c = 0; c < 9; ++c) {
      uint8_t o = s.cheese_outcomes[c];
      bool was_cheese = (c == 1 || c == 7 || c == 4);
      if (!was_cheese) CHECK(o == 2);
      if (o == 0 || o == 3) collected += 1.0f;
      if (o == 1) collected += 1.0f;
    }
    NEAR(collected, s.final_p1_score + s.final_p2_score, 1e-6);
    CHECK(s.result == (s.final_p1_score > s.final_p2_score ? 1 : s.final_p2_score > s.final_p1_score ? 2 : 0));
  }

  std::printf("%d checks passed, %d failed\n", g_pass, g_fail);
  return g_fail == 0 ? 0 : 1;
}
