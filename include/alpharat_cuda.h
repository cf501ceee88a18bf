/*
 * alpharat_cuda.h — C-ABI of libalpharat_cuda.so, the B200-native `backend: cuda`
 * for the batched self-play MCTS path of mintiti/alpharat.
 *
 * Every entry point replaces one PyO3 binding of the reference (paths relative to the
 * reference tree); INTEGRATION.md shows the ctypes stub a maintainer adds on the
 * reference side.
 *
 *   ar_search_batch   <- rust_mcts_search      crates/alpharat-mcts/src/bindings.rs:228-304
 *   ar_selfplay_run   <- rust_self_play        crates/alpharat-sampling/src/bindings.rs:268-483
 *   ar_engine_load_weights <- OnnxBackend::with_provider / TensorrtBackend::new
 *                                              crates/alpharat-sampling/src/bindings.rs:222-256
 *   ar_encode_observations <- FlatEncoder::encode_into
 *                                              crates/alpharat-sampling/src/flat_encoder.rs:52-124
 *   ar_nn_forward     <- Backend::evaluate_batch (ONNX / TensorRT)
 *                                              crates/alpharat-sampling/src/backends/onnx.rs:176-245
 *   ar_engine_set_eval_cache <- CachedBackend::new (rust_self_play's cache_size)
 *                                              crates/alpharat-sampling/src/cached_backend.rs:54-120
 *
 * Plain C types only: pointers and sizes, caller-owned host buffers, fixed-layout
 * little-endian PODs.  The engine owns all device memory.  An engine handle is
 * single-threaded (one handle per GPU, one host thread per handle); the progress block
 * may be read concurrently from other threads.
 */
#ifndef ALPHARAT_CUDA_H
#define ALPHARAT_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AR_ABI_VERSION 2
#define AR_MAX_CELLS 256 /* POD capacity (16x16); kernels in this build accept <= 64 cells */
#define AR_NUM_ACTIONS 5 /* UP=0 RIGHT=1 DOWN=2 LEFT=3 STAY=4 (pyrat_engine Direction)      */

typedef enum ar_status {
  AR_OK = 0,
  AR_ERR_INVALID_ARG = 1, /* maps to ValueError                                         */
  AR_ERR_CUDA = 2,        /* maps to RuntimeError (BackendError in the reference)       */
  AR_ERR_POOL_OVERFLOW = 3, /* node pool / depth stack exhausted; raise pool_nodes      */
  AR_ERR_NONFINITE = 4,   /* NN produced a non-finite value (onnx.rs:233-241)           */
  AR_ERR_UNSUPPORTED = 5, /* board larger than this build supports, unknown arch ...    */
  AR_ERR_NO_WEIGHTS = 6
} ar_status;

/* One PyRat position plus its maze.  Stands in for pyrat::GameState where the reference
 * passes a PyRat object across PyO3 (mcts/bindings.rs:250) or builds games inside Rust
 * (sampling/bindings.rs:489-533).  cell = y*width + x, Y-up (origin bottom-left). */
typedef struct ar_game_pod {
  uint8_t width, height;
  uint8_t p1_x, p1_y, p2_x, p2_y;
  uint8_t p1_mud, p2_mud;   /* mud_timer: turns the player is still stuck            */
  uint16_t turn, max_turns;
  uint16_t reserved0, reserved1;
  float p1_score, p2_score; /* multiples of 0.5                                       */
  /* move_cost[cell*4+dir]: 0 = wall/boundary, 1 = open, c>=2 = mud of cost c           */
  uint8_t move_cost[AR_MAX_CELLS * 4];
  uint8_t cheese[AR_MAX_CELLS / 8]; /* bit (cell&7) of byte cell>>3                    */
} ar_game_pod;

/* SearchConfig (crates/alpharat-mcts/src/search.rs:18-58) + n_sims/batch_size of
 * run_search (search.rs:362-370). */
typedef struct ar_search_cfg {
  uint32_t simulations;
  uint32_t batch_size;
  float c_puct;
  float fpu_reduction;
  float force_k;
  float noise_epsilon;
  float noise_concentration;
  uint32_t collision_limit_min;
  uint32_t collision_limit_max;
  uint32_t collision_scaling_start;
  uint32_t collision_scaling_end;
  float collision_scaling_power;
} ar_search_cfg;

/* SearchResult (search.rs:304-325) plus the raw integer root edge visits (expand_visits,
 * node.rs:216-223) that the reference only exposes inside Rust. */
typedef struct ar_search_result {
  float policy_p1[5], policy_p2[5];
  float value_p1, value_p2;
  float visit_counts_p1[5], visit_counts_p2[5]; /* pruned, f32 */
  float prior_p1[5], prior_p2[5];
  uint32_t total_visits, nn_evals, terminals, collisions;
  uint32_t raw_visits_p1[5], raw_visits_p2[5];
  uint32_t node_count; /* MCTSTree::node_count after the search (tree.rs:267) */
  uint32_t reserved;
} ar_search_result;

/* PositionRecord (crates/alpharat-sampling/src/selfplay.rs:82-102) plus the search
 * counters of that move (parity diagnostics). */
typedef struct ar_position_record {
  uint8_t p1_x, p1_y, p2_x, p2_y;
  uint8_t p1_mud, p2_mud;
  uint8_t action_p1, action_p2;
  uint16_t turn, reserved;
  float p1_score, p2_score;
  ar_search_result search;
  uint8_t cheese[AR_MAX_CELLS / 8];
} ar_position_record;

/* GameRecord minus the per-position array (selfplay.rs:106-132). */
typedef struct ar_game_summary {
  uint32_t game_index;
  uint32_t n_positions;
  float final_p1_score, final_p2_score;
  uint8_t result; /* 0 draw, 1 P1 win, 2 P2 win (GameOutcome, selfplay.rs:62-66) */
  uint8_t reserved[3];
  uint16_t cheese_available, reserved1;
  uint64_t total_simulations, total_nn_evals, total_terminals, total_collisions;
  /* CheeseOutcome per cell: 0 P1, 1 simultaneous, 2 uncollected, 3 P2 (selfplay.rs:73-78,415-471) */
  uint8_t cheese_outcomes[AR_MAX_CELLS];
} ar_game_summary;

/* SelfPlayStats (selfplay.rs:136-158) + roofline counters. */
typedef struct ar_stats {
  uint32_t total_games;
  uint32_t p1_wins, p2_wins, draws;
  uint64_t total_positions, total_simulations;
  uint64_t total_nn_evals, total_terminals, total_collisions;
  uint64_t cache_hits, cache_misses;
  double elapsed_secs;
  float total_cheese_collected;
  uint32_t total_cheese_available;
  uint32_t min_turns, max_turns;
  /* device-side measurements of the same run */
  double device_ms;          /* CUDA-event time of all kernels of the run            */
  uint64_t path_nodes;       /* node visits (select+backup pairs) — 288 B each         */
  uint64_t new_nodes;        /* nodes created — 240 B each                             */
  uint64_t kernel_launches;  /* kernels launched by this call                          */
  uint64_t h2d_bytes, d2h_bytes;
} ar_stats;

/* SelfPlayProgress (selfplay.rs:343-348): written by the engine, read concurrently. */
typedef struct ar_progress {
  volatile uint32_t games_completed;
  volatile uint32_t reserved;
  volatile uint64_t positions_completed;
  volatile uint64_t simulations_completed;
  volatile uint64_t nn_evals_completed;
} ar_progress;

typedef struct ar_engine_cfg {
  uint32_t abi_version;      /* AR_ABI_VERSION                                         */
  int32_t device;            /* CUDA ordinal                                           */
  uint32_t concurrent_games; /* resident game trees                                    */
  uint32_t pool_nodes;       /* node-pool capacity per tree (<= 65535), 0 = auto       */
  uint32_t max_cells;        /* width*height upper bound for this engine               */
  uint32_t max_turns;        /* upper bound on max_turns (depth stack)                 */
  uint32_t max_batch_size;   /* upper bound on ar_search_cfg.batch_size                */
  uint32_t max_simulations;  /* used for pool auto-sizing                              */
  uint32_t tree_engine;      /* uniform-prior runs: AR_TREE_WARP (default), AR_TREE_THREAD, AR_TREE_HALF */
} ar_engine_cfg;

/* Two device engines play the uniform-prior path, bit-identically (both are checked against the
 * oracle).  AR_TREE_WARP: one warp per game tree, fixed node pools, lowest latency per game.
 * AR_TREE_THREAD: one thread per game tree (32 trees per warp), paged node pools shared by all
 * resident trees; measured slower on B200 (profiles/r2_summary.md), kept as a selectable engine. */
#define AR_TREE_WARP 0
#define AR_TREE_THREAD 1
#define AR_TREE_HALF 2 /* two trees per warp (16 lanes each); same records, pools and results as AR_TREE_WARP */

/* One named f32 tensor of a torch state_dict (host memory). */
typedef struct ar_tensor_desc {
  const char* name;
  const float* data;
  int32_t ndim;
  int64_t shape[4];
} ar_tensor_desc;

typedef enum ar_arch {
  AR_ARCH_UNIFORM = 0, /* SmartUniformBackend, backend.rs:92-103 */
  AR_ARCH_MLP = 1,     /* alpharat/nn/models/mlp.py:120-153       */
  AR_ARCH_SYMMETRIC = 2, /* alpharat/nn/models/symmetric.py:124-207 */
  AR_ARCH_CNN = 3      /* alpharat/nn/models/cnn/model.py:117-220 */
} ar_arch;

typedef struct ar_engine ar_engine;

ar_status ar_engine_create(const ar_engine_cfg* cfg, ar_engine** out);
void ar_engine_destroy(ar_engine* e);
const char* ar_last_error(const ar_engine* e); /* NULL engine -> last create() error */
uint32_t ar_abi_version(void);

/* Load an evaluator.  arch == AR_ARCH_UNIFORM clears it.  Tensors are the model's
 * state_dict in f32; the engine folds eval-mode BatchNorm, converts and uploads. */
ar_status ar_engine_load_weights(ar_engine* e, int32_t arch, int32_t width, int32_t height,
                                 const ar_tensor_desc* tensors, int32_t n_tensors);

/* Evaluation cache of the NN-guided mode: entries_per_tree positions per resident tree (rounded up to a
 * power of two, at most 65536; 0 disables).  Replaces the thread-local NNCache behind CachedBackend: a leaf
 * whose exact position was scored before skips the evaluator; search results do not depend on it.
 * ar_stats.cache_hits / cache_misses report the lookups of a run. */
ar_status ar_engine_set_eval_cache(ar_engine* e, uint32_t entries_per_tree);

/* n independent fresh-tree searches (rust_mcts_search semantics, one RNG per search
 * seeded with SmallRng::seed_from_u64(seeds[i])). */
ar_status ar_search_batch(ar_engine* e, const ar_game_pod* games, int32_t n,
                          const ar_search_cfg* cfg, const uint64_t* seeds,
                          ar_search_result* out);

/* Play n games to completion (play_game, selfplay.rs:515-598), one RNG per game seeded
 * with seeds[i].  positions has n * positions_stride entries (stride >= max_turns of
 * every game).  progress and stats may be NULL. */
ar_status ar_selfplay_run(ar_engine* e, const ar_game_pod* games, int32_t n,
                          const ar_search_cfg* cfg, const uint64_t* seeds,
                          ar_game_summary* summaries, ar_position_record* positions,
                          int32_t positions_stride, ar_progress* progress, ar_stats* stats);

/* Device-resident variant used for kernel-only timing: upload once, run many times. */
ar_status ar_selfplay_upload(ar_engine* e, const ar_game_pod* games, int32_t n,
                             const uint64_t* seeds);
ar_status ar_selfplay_run_resident(ar_engine* e, const ar_search_cfg* cfg, ar_stats* stats);
ar_status ar_selfplay_download(ar_engine* e, ar_game_summary* summaries,
                               ar_position_record* positions, int32_t positions_stride);

/* The resident batch's records packed on the DEVICE (game-major, n_positions records per game) and its
 * summaries: device pointers, valid until the next call on this engine.  For device-to-device gathers of
 * recorded batches across GPUs (the aggregation of run_self_play, selfplay.rs:657-703, without a host
 * bounce).  `summaries` (host, n entries) receives the summaries too: game lengths give the offsets. */
ar_status ar_selfplay_pack_device(ar_engine* e, ar_game_summary* summaries, const void** d_summaries,
                                  const void** d_records, uint64_t* n_records);

/* Streaming self-play (continuous game feed).  The reference's worker pool never idles between batches:
 * threads keep claiming game indices until the run is over (game_worker_loop, selfplay.rs:609-650,
 * run_self_play_to_disk 706-808).  A lone GPU batch instead ends with a tail in which its last, longest
 * games run alone; with several batches in flight the next batch's blocks take over the SMs as the previous
 * batch's blocks finish, so throughput does not depend on the batch size.  `n_buffers` batches of up to
 * `max_games` games can be in flight; submit / collect a buffer in turn.  Uniform-prior AR_TREE_WARP only.
 *   ar_stream_submit   host games + seeds -> pinned staging -> device (async) and launch (async)
 *   ar_stream_collect  wait for that buffer, download its records (same layout as ar_selfplay_run)
 *   ar_stream_launch / ar_stream_wait   re-run a buffer whose inputs are already resident (kernel-only timing)
 *   ar_stream_submit with cfg == NULL   upload only
 *   ar_stream_times    start / end of a buffer's last completed launch on the device clock (ms since open)
 *   ar_stream_elapsed_ms   device time from the start of buffer `first`'s launch to the end of `last`'s */
ar_status ar_stream_open(ar_engine* e, int32_t n_buffers, int32_t max_games, int32_t positions_stride);
void ar_stream_close(ar_engine* e);
ar_status ar_stream_submit(ar_engine* e, int32_t buffer, const ar_game_pod* games, int32_t n,
                           const ar_search_cfg* cfg, const uint64_t* seeds);
ar_status ar_stream_collect(ar_engine* e, int32_t buffer, ar_game_summary* summaries,
                            ar_position_record* positions, int32_t positions_stride, ar_stats* stats);
ar_status ar_stream_launch(ar_engine* e, int32_t buffer, const ar_search_cfg* cfg);
ar_status ar_stream_wait(ar_engine* e, int32_t buffer, ar_stats* stats);
ar_status ar_stream_times(ar_engine* e, int32_t buffer, double* start_ms, double* end_ms);
ar_status ar_stream_elapsed_ms(ar_engine* e, int32_t first, int32_t last, double* ms);

/* FlatEncoder on device: obs is n * (7*w*h+6) f32, host memory. */
ar_status ar_encode_observations(ar_engine* e, const ar_game_pod* games, int32_t n, float* obs);

/* Leaf evaluator on device for n positions: policy_p1/p2 are n*5, values n (host). */
ar_status ar_nn_forward(ar_engine* e, const ar_game_pod* games, int32_t n, float* policy_p1,
                        float* policy_p2, float* value_p1, float* value_p2);

#ifdef __cplusplus
}
#endif
#endif /* ALPHARAT_CUDA_H */
