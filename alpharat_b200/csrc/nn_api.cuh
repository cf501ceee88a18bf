// nn_api.cuh — interface between the tree engine (engine.cu) and the leaf evaluators
// (nn_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/alpharat_cuda.h"

namespace ar {

// One leaf to evaluate: the mutable part of the game state + which game's maze it belongs to.
struct __align__(16) EvalRow {
  uint64_t cheese;
  uint32_t pos;      // p1 | p2 << 8 | mud1 << 16 | mud2 << 24
  uint32_t score;    // s1x2 | s2x2 << 16
  uint32_t game_idx;
  uint16_t turn, max_turns;
  uint32_t pad[2];
};
static_assert(sizeof(EvalRow) == 32, "EvalRow layout");

struct MlpWeights {
  const uint8_t* w1;  // k1_blocks x [256 x 64] bf16, 128B-swizzled images
  const uint8_t* w2;  // 4 x [256 x 64]
  const uint8_t* w3;  // 4 x [16 x 64]   rows: policy_p1[5], policy_p2[5], value[2], 0[4]
  const float* b1;
  const float* b2;
  const float* b3;
  int k1_blocks;
};

// A loaded leaf evaluator (replaces `dyn Backend`, crates/alpharat-mcts/src/backend.rs:75-82, for the
// NN-guided mode).  forward() scores `rows` into out[row][12] = policy_p1[5], policy_p2[5], v1, v2.
// n_rows_dev != nullptr: the row count is read on device (the self-play loop never syncs).
// maze_tab: build_maze_table() of `games`.
struct LeafEvaluator {
  virtual ~LeafEvaluator() {}
  virtual int load(const ar_tensor_desc* tensors, int n, int width, int height, std::string& err) = 0;
  virtual cudaError_t forward(const EvalRow* rows, const uint32_t* n_rows_dev, int n_rows_max,
                              const ar_game_pod* games, const uint16_t* maze_tab, float* out,
                              int* error_flag, cudaStream_t stream) const = 0;
};

struct MlpModel : LeafEvaluator {
  uint8_t *d_w1 = nullptr, *d_w2 = nullptr, *d_w3 = nullptr;
  float* d_b = nullptr;
  int k1_blocks = 0, obs_dim = 0, n_sms = 148;
  size_t smem_bytes = 0;
  bool loaded = false;
  ~MlpModel() override { release(); }
  int load(const ar_tensor_desc* tensors, int n, int width, int height, std::string& err) override;
  cudaError_t forward(const EvalRow* rows, const uint32_t* n_rows_dev, int n_rows_max,
                      const ar_game_pod* games, const uint16_t* maze_tab, float* out, int* error_flag,
                      cudaStream_t stream) const override;
  void release();
};

LeafEvaluator* make_symmetric_evaluator();  // nn_symmetric.cu
LeafEvaluator* make_cnn_evaluator();        // nn_cnn.cu

// Per-game maze channels of the observation as bf16 (flat_encoder.rs:62-80): out[game][cell * 4 + dir],
// 256 entries per game, zero past the board.  The maze never changes during a game, so the
// evaluators copy these rows instead of re-deriving 4 * cells values for every leaf.
constexpr int MAZE_TAB_STRIDE = 256;
cudaError_t build_maze_table(const ar_game_pod* games, int n, uint16_t* out, cudaStream_t stream);

cudaError_t encode_f32(const EvalRow* rows, int n, const ar_game_pod* games, int obs_dim, float* out,
                       cudaStream_t stream);

}  // namespace ar
