// tt_emul.cpp — TEST INFRASTRUCTURE ONLY.
//
// Compiles the device core (alpharat_b200/csrc/tree_thread.cuh) for the host and runs T logical
// "threads" round-robin, one tt_step each per sweep, exactly as the warps of the CUDA kernel
// interleave them.  It exists so that the state machine, the paged pools and the tree compaction
// can be checked bit-for-bit against oracle/ in the CPU-only container; the product never loads
// it (alpharat_b200/ has no reference to tests/).  Build: g++ -O2 -ffp-contract=off -shared.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../alpharat_b200/csrc/host_tables.hpp"
#include "../../alpharat_b200/csrc/tree_thread.cuh"

template <int NW>
static int run_nw(const ar_game_pod* games, int n, const ar_search_cfg* cfg, const uint64_t* seeds,
                  int n_threads, int n_pages, int search_only, ar_game_summary* summaries,
                  ar_position_record* positions, int stride, ar_search_result* search_out,
                  unsigned long long* counters);

extern "C" int tt_emul_run(const ar_game_pod* games, int n, const ar_search_cfg* cfg, const uint64_t* seeds,
                           int n_threads, int n_pages, int search_only, ar_game_summary* summaries,
                           ar_position_record* positions, int stride, ar_search_result* search_out,
                           unsigned long long* counters /* [4]: path_nodes new_nodes peak_pages steps */) {
  bool big = false;  // boards over 64 cells need the 4-word cheese bitboard
  for (int i = 0; i < n; ++i) big = big || (int)games[i].width * games[i].height > 64;
  return big ? run_nw<4>(games, n, cfg, seeds, n_threads, n_pages, search_only, summaries, positions, stride, search_out, counters)
             : run_nw<1>(games, n, cfg, seeds, n_threads, n_pages, search_only, summaries, positions, stride, search_out, counters);
}

template <int NW>
static int run_nw(const ar_game_pod* games, int n, const ar_search_cfg* cfg, const uint64_t* seeds,
                  int n_threads, int n_pages, int search_only, ar_game_summary* summaries,
                  ar_position_record* positions, int stride, ar_search_result* search_out,
                  unsigned long long* counters) {
  using namespace tt;
  using T = TT<NW>;
  using TState = typename T::TState;
  using TArr = typename T::TArr;
  if (n_threads < 1 || n_pages < n_threads) return -1;
  Ctx c;
  memset(&c, 0, sizeof(c));
  std::vector<uint8_t> arena((size_t)n_pages * PAGE_BYTES + 16);
  uint8_t* abase = arena.data();
  abase += (16 - ((uintptr_t)abase & 15)) & 15;
  c.arena = abase;
  c.n_pages = (uint32_t)n_pages;
  c.bitmap_words = ((uint32_t)n_pages + 31) / 32;
  std::vector<uint32_t> bitmap(c.bitmap_words, 0);
  for (int i = 0; i < n_threads; ++i) bitmap[i >> 5] |= 1u << (i & 31);
  c.page_bitmap = bitmap.data();
  c.pt_stride = 128;
  std::vector<uint32_t> pts((size_t)n_threads * c.pt_stride, 0);
  c.page_tables = pts.data();
  const uint32_t coll_len = 1u << 17;
  std::vector<uint16_t> coll = ar_host::collision_table(*cfg, coll_len);
  c.coll_table = coll.data();
  c.coll_len = coll_len;
  c.sp.c_puct = cfg->c_puct; c.sp.fpu_reduction = cfg->fpu_reduction; c.sp.force_k = cfg->force_k;
  c.sp.noise_epsilon = cfg->noise_epsilon; c.sp.noise_concentration = cfg->noise_concentration;
  c.sp.n_sims = cfg->simulations; c.sp.batch_size = cfg->batch_size;
  c.games = games; c.seeds = seeds; c.n_games = n;
  uint32_t next_game = 0;
  c.next_game = &next_game;
  c.summaries = summaries; c.positions = positions; c.pos_stride = stride;
  c.search_out = search_out; c.search_only = search_only;
  unsigned long long ctr[8] = {0};
  c.counters = ctr;
  int err = 0;
  c.error_flag = &err;
  c.progress = nullptr;

  std::vector<uint32_t> maze((size_t)T::MAZE_WORDS * n_threads, 0);
  std::vector<TState> st(n_threads);
  std::vector<TArr> arr(n_threads);
  for (int t = 0; t < n_threads; ++t) T::tt_init(st[t], c, (uint32_t)t, maze.data() + t, n_threads);
  unsigned long long steps = 0, peak = 0;
  for (;;) {
    int alive = 0;
    for (int t = 0; t < n_threads; ++t) {
      if (st[t].phase == PH_EXIT) continue;
      alive += 1;
      T::template tt_step<false>(st[t], arr[t], c);
      steps += 1;
    }
    if ((steps & 0xfff) < (unsigned)n_threads) {
      unsigned long long used = 0;
      for (uint32_t w : bitmap) used += (unsigned long long)__builtin_popcount(w);
      peak = used > peak ? used : peak;
    }
    if (!alive) break;
  }
  int rc = 0;
  for (int t = 0; t < n_threads; ++t) {
    ctr[0] += st[t].path_nodes;
    ctr[1] += st[t].new_nodes;
    if (st[t].error) rc = (int)st[t].error;
  }
  // every page except the threads' own first pages must have been returned
  unsigned long long used = 0;
  for (uint32_t w : bitmap) used += (unsigned long long)__builtin_popcount(w);
  if (rc == 0 && used != (unsigned long long)n_threads) rc = -100;
  if (counters) { counters[0] = ctr[0]; counters[1] = ctr[1]; counters[2] = peak; counters[3] = steps; }
  return rc;
}
