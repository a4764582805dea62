// Job table of the tensor-core weight-gradient kernel (tc_wgrad.cu).
#pragma once
#include "common.cuh"

namespace agx {
namespace tc {

constexpr int WG_MAX_JOBS = 24;

struct WgJob {
  const float* dY;      // [M][FP] upstream gradient
  const float* mask;    // optional [M][FP]: dY is multiplied by (mask > 0)   (ReLU backward)
  const float* X;       // [M][ldx] layer input
  float* dW;            // reference-layout weight gradient, row stride ld; columns col0 .. col0 + K - 1 receive the product
  float* db;            // optional bias gradient
  int64_t M;
  const int32_t* m_limit;  // optional device scalar: only the first min(M, *m_limit) rows exist (relations actually built)
  int ldx, kx;          // kx: readable X columns (multiple of 4)
  int npad;             // MMA N extent: 16, 32 or FP; X column npad - 1 is replaced by the constant 1 when bias_col
  int bias_col;
  int ld, col0, F, K;
  int chain_head, next; // jobs adding into the same destination form a chain (next = -1 ends it); only the head reduces
};
struct WgArgs {
  WgJob job[WG_MAX_JOBS];
  int njobs;
  float* part;          // [total CTAs][FP * FP] partial products
};

}  // namespace tc

size_t tc_wgrad_part_floats(int max_ctas);
// Launches the batch (gradient kernel + fixed-order reduction).  `part` holds max_ctas partial products.
int tc_wgrad_batch(cudaStream_t st, tc::WgArgs& a, float* part, int max_ctas);

}  // namespace agx
