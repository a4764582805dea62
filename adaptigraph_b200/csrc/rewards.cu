// MPC reward tail in ONE kernel: everything the planner evaluates after the last rollout step.
//
//   running_cost                       src/planning/plan.py:27-59
//   chamfer / box_loss  (error term)   src/planning/losses.py:4-10, 25-35      (plan.py:146, :155 pick one)
//   rope / cloth / granular penalty    src/planning/losses.py:37-92            (plan.py:160-165 pick one)
//
// The reference evaluates these as ~60 eager tensor ops per call (materialising (B, M, N, 3) tensors for the chamfer distance) and
// synchronises twice (`error.max().item()`, `action_state_max_dist.max().item()`).  Here one CTA per (sample, look-ahead step)
// keeps the step's particles in shared memory and produces the three terms of that cell; the two batch-wide maxima go through
// device atomics and the last CTA to finish (ticket) turns the cells into `reward_seqs` -- one launch, no host round trip.
// Arithmetic is the reference's fp32 (norms as sqrt of unfused sums of squares, exp(-100 d)); only the order of the means over
// particles differs from torch's.
#include "common.cuh"

namespace agx {

constexpr int RW_THREADS = 256;

__device__ __forceinline__ float rw_block_reduce(float v, float* red, int op /*0 sum, 1 min, 2 max*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float u = __shfl_xor_sync(0xffffffffu, v, o);
    v = op == 0 ? v + u : (op == 1 ? fminf(v, u) : fmaxf(v, u));
  }
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = red[0];
  for (int w = 1; w < RW_THREADS / 32; ++w) t = op == 0 ? t + red[w] : (op == 1 ? fminf(t, red[w]) : fmaxf(t, red[w]));   // fixed order
  return t;
}

__device__ __forceinline__ float norm2(float dx, float dz) {   // torch.norm(., dim=-1) on 2 components
  return __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dz, dz)));
}

struct RewardArgs {
  const float* state;      // (bsz, L, n, 3) predicted particles
  const float* action;     // (bsz, L, action_dim)
  const float* state_cur;  // (n, 3)
  const float* bbox;       // (2, 2)
  const float* target;     // error_mode 0: (M, 3) target points; 1: (2, 2) target box
  int bsz, L, n, action_dim, M, error_mode, penalty_mode;
  float sim_real_ratio;
  float* cells;            // workspace (bsz * L, 4): error, penalty (rope / granular) or clamped min distance (cloth), clamped max distance (cloth), box penalty
  int* scratch;            // workspace [4]: int view of max error, of max clamped max-distance, ticket, -
  float* reward;           // (bsz)
};

__global__ void __launch_bounds__(RW_THREADS) running_cost_kernel(const RewardArgs a) {
  extern __shared__ float rw_smem[];
  __shared__ float red[RW_THREADS / 32];
  __shared__ int last_s;
  float* xs = rw_smem;                 // [n][3] this cell's particles
  float* ys = rw_smem + 3 * a.n;       // [M][3] target points (chamfer)
  const int cell = blockIdx.x, b = cell / a.L, l = cell - b * a.L, tid = threadIdx.x, n = a.n;
  const float* xb = a.state + (size_t)cell * n * 3;
  for (int i = tid; i < 3 * n; i += RW_THREADS) xs[i] = xb[i];
  if (a.error_mode == 0)
    for (int i = tid; i < 3 * a.M; i += RW_THREADS) ys[i] = a.target[i];
  __syncthreads();
  const float INF = __int_as_float(0x7f800000);

  // ---- error term of the cell
  float error;
  if (a.error_mode == 0) {             // chamfer (losses.py:4-10): mean_m min_n |x_n - y_m| + mean_n min_m |x_n - y_m|
    float sum_m = 0.f, sum_n = 0.f;
    for (int m = tid; m < a.M; m += RW_THREADS) {
      const float a0 = ys[3 * m], a1 = ys[3 * m + 1], a2 = ys[3 * m + 2];
      float best = INF;
      for (int i = 0; i < n; ++i) {
        const float d0 = xs[3 * i] - a0, d1 = xs[3 * i + 1] - a1, d2 = xs[3 * i + 2] - a2;
        best = fminf(best, __fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)));
      }
      sum_m += __fsqrt_rn(best);
    }
    for (int i = tid; i < n; i += RW_THREADS) {
      const float a0 = xs[3 * i], a1 = xs[3 * i + 1], a2 = xs[3 * i + 2];
      float best = INF;
      for (int m = 0; m < a.M; ++m) {
        const float d0 = a0 - ys[3 * m], d1 = a1 - ys[3 * m + 1], d2 = a2 - ys[3 * m + 2];
        best = fminf(best, __fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)));
      }
      sum_n += __fsqrt_rn(best);
    }
    const float tm = rw_block_reduce(sum_m, red, 0), tn = rw_block_reduce(sum_n, red, 0);
    error = tm / (float)a.M + tn / (float)n;
  } else {                             // box_loss (losses.py:25-35)
    const float xmin = a.target[0], xmax = a.target[1], zmin = a.target[2], zmax = a.target[3];
    float s = 0.f;
    for (int i = tid; i < n; i += RW_THREADS) {
      const float x = xs[3 * i], z = xs[3 * i + 2];
      const float xd = __fadd_rn(fmaxf(__fsub_rn(xmin, x), 0.f), fmaxf(__fsub_rn(x, xmax), 0.f));
      const float zd = __fadd_rn(fmaxf(__fsub_rn(zmin, z), 0.f), fmaxf(__fsub_rn(z, zmax), 0.f));
      s += __fsqrt_rn(__fadd_rn(__fmul_rn(xd, xd), __fmul_rn(zd, zd)));
    }
    error = rw_block_reduce(s, red, 0) / (float)n;
  }

  // ---- collision penalty of the cell: distance of the pusher's start point(s) to the particles BEFORE this push
  // (the current state for l = 0, the previous prediction afterwards: losses.py:43-44, :87-88; cloth always uses the current state)
  const float* act = a.action + (size_t)cell * a.action_dim;
  const float ax = act[0], az = act[1];
  const float* prev = (l == 0 || a.penalty_mode == 1) ? a.state_cur : a.state + (size_t)(cell - 1) * n * 3;
  float pen = 0.f, dmax_c = 0.f;
  if (a.penalty_mode == 1) {           // cloth_penalty (losses.py:51-65)
    float dmin = INF, dmax = 0.f;
    for (int i = tid; i < n; i += RW_THREADS) {
      const float d = norm2(__fsub_rn(ax, prev[3 * i]), __fsub_rn(az, prev[3 * i + 2]));
      dmin = fminf(dmin, d);
      dmax = fmaxf(dmax, d);
    }
    dmin = rw_block_reduce(dmin, red, 1);
    dmax = rw_block_reduce(dmax, red, 2);
    pen = fmaxf(__fsub_rn(dmin, (float)(0.005 * (double)a.sim_real_ratio)), 0.f);        // finished by the last CTA (needs the batch maximum)
    dmax_c = fminf(dmax, (float)(0.4 * (double)a.sim_real_ratio));
  } else {
    float dmin = INF;
    if (a.penalty_mode == 0) {         // rope_penalty (losses.py:37-49)
      for (int i = tid; i < n; i += RW_THREADS)
        dmin = fminf(dmin, norm2(__fsub_rn(ax, prev[3 * i]), __fsub_rn(az, prev[3 * i + 2])));
    } else {                           // granular_penalty (losses.py:67-92): 9 points along the pusher blade
      const float pr = (float)(0.05 * (double)a.sim_real_ratio);   // Python float products in the reference, rounded once
      const float dx = __fmul_rn(pr, sinf(act[2])), dz = __fmul_rn(-pr, cosf(act[2]));
      const float f[9] = {-1.f, -0.75f, -0.5f, -0.25f, 0.f, 0.25f, 0.5f, 0.75f, 1.f};
      float px[9], pz[9];
#pragma unroll
      for (int q = 0; q < 9; ++q) {
        // the reference writes x - 0.75 * dx etc.: the product first, then the sum (no contraction); +-1 and 0 have no product
        px[q] = q == 4 ? ax : (f[q] < 0.f ? __fsub_rn(ax, f[q] == -1.f ? dx : __fmul_rn(-f[q], dx)) : __fadd_rn(ax, f[q] == 1.f ? dx : __fmul_rn(f[q], dx)));
        pz[q] = q == 4 ? az : (f[q] < 0.f ? __fsub_rn(az, f[q] == -1.f ? dz : __fmul_rn(-f[q], dz)) : __fadd_rn(az, f[q] == 1.f ? dz : __fmul_rn(f[q], dz)));
      }
      for (int i = tid; i < n; i += RW_THREADS) {
        const float x = prev[3 * i], z = prev[3 * i + 2];
#pragma unroll
        for (int q = 0; q < 9; ++q) dmin = fminf(dmin, norm2(__fsub_rn(px[q], x), __fsub_rn(pz[q], z)));
      }
    }
    dmin = rw_block_reduce(dmin, red, 1);
    const float d = fmaxf(__fsub_rn(dmin, (float)(0.02 * (double)a.sim_real_ratio)), 0.f);
    pen = expf(__fmul_rn(-d, 100.f));
  }

  // ---- box penalty of the cell (plan.py:41-51): how far the particles' bounding box reaches towards / past the workspace box
  float xmn = INF, xmx = -INF, zmn = INF, zmx = -INF;
  for (int i = tid; i < n; i += RW_THREADS) {
    const float x = xs[3 * i], z = xs[3 * i + 2];
    xmn = fminf(xmn, x); xmx = fmaxf(xmx, x); zmn = fminf(zmn, z); zmx = fmaxf(zmx, z);
  }
  xmn = rw_block_reduce(xmn, red, 1); xmx = rw_block_reduce(xmx, red, 2);
  zmn = rw_block_reduce(zmn, red, 1); zmx = rw_block_reduce(zmx, red, 2);
  float boxpen = 0.f;
  if (tid == 0) {
    const float p0 = fmaxf(__fsub_rn(xmn, a.bbox[0]), 0.f), p1 = fmaxf(__fsub_rn(a.bbox[1], xmx), 0.f);
    const float p2 = fmaxf(__fsub_rn(zmn, a.bbox[2]), 0.f), p3 = fmaxf(__fsub_rn(a.bbox[3], zmx), 0.f);
    boxpen = fmaxf(fmaxf(expf(__fmul_rn(-p0, 100.f)), expf(__fmul_rn(-p1, 100.f))), fmaxf(expf(__fmul_rn(-p2, 100.f)), expf(__fmul_rn(-p3, 100.f))));
    float4* c = reinterpret_cast<float4*>(a.cells) + cell;
    *c = make_float4(error, pen, dmax_c, boxpen);
    atomicMax(&a.scratch[0], __float_as_int(error));      // error, dmax_c >= 0: the int view orders like the floats
    atomicMax(&a.scratch[1], __float_as_int(dmax_c));
    __threadfence();
    last_s = atomicAdd(&a.scratch[2], 1) == (int)gridDim.x - 1;
  }
  __syncthreads();
  if (!last_s) return;

  // ---- the last CTA: reward_seqs (plan.py:37, :53)
  __threadfence();
  const float emax = __int_as_float(__ldcg(&a.scratch[0])), gmax = __int_as_float(__ldcg(&a.scratch[1]));
  const float weight = (float)(2.0 / ((double)emax + 1e-6));              // Python float arithmetic in the reference
  const float4* cells = reinterpret_cast<const float4*>(a.cells);
  for (int s = tid; s < a.bsz; s += RW_THREADS) {
    float psum = 0.f, bsum = 0.f;
    for (int q = 0; q < a.L; ++q) {
      const float4 c = __ldcg(cells + (size_t)s * a.L + q);
      float p = c.y;
      if (a.penalty_mode == 1) {
        p = __fsub_rn(__fsub_rn(1.f, expf(__fmul_rn(-c.y, 100.f))), __fmul_rn(__fdiv_rn(c.z, gmax), 0.2f));   // losses.py:63-64
        a.cells[((size_t)s * a.L + q) * 4 + 1] = p;       // the finished term replaces its raw distance (read by rewards.cloth_penalty)
      }
      psum += p;
      bsum += c.w;
    }
    const float e_last = __ldcg(cells + (size_t)s * a.L + (a.L - 1)).x;
    a.reward[s] = __fsub_rn(__fsub_rn(__fmul_rn(-weight, e_last), __fmul_rn(5.f, psum / (float)a.L)), __fmul_rn(5.f, bsum / (float)a.L));
  }
  if (tid == 0) { a.scratch[0] = 0; a.scratch[1] = 0; a.scratch[2] = 0; }   // ready for the next call on this workspace
}

}  // namespace agx

extern "C" {

size_t agx_running_cost_workspace_bytes(int32_t bsz, int32_t L) {
  if (bsz <= 0 || L <= 0) return 0;
  return agx::align_up((size_t)bsz * L * 16, 256) + 256;
}

int agx_running_cost(const float* state, const float* action, int32_t action_dim, const float* state_cur, const float* bbox,
                     int32_t error_mode, const float* target, int32_t M, int32_t penalty_mode, float sim_real_ratio, int32_t bsz,
                     int32_t L, int32_t n, void* workspace, size_t workspace_bytes, float* reward, agx_stream_t stream) {
  using namespace agx;
  AGX_REQUIRE(state && action && state_cur && bbox && target && reward, AGX_ERR_ARG, "running_cost: null pointer argument");
  AGX_REQUIRE(bsz > 0 && L > 0 && n > 0 && action_dim >= 3, AGX_ERR_ARG, "running_cost: bsz=%d L=%d n=%d action_dim=%d", bsz, L, n, action_dim);
  AGX_REQUIRE(error_mode == AGX_ERROR_CHAMFER || error_mode == AGX_ERROR_BOX, AGX_ERR_ARG, "running_cost: bad error_mode %d", error_mode);
  AGX_REQUIRE(penalty_mode >= AGX_PENALTY_ROPE && penalty_mode <= AGX_PENALTY_GRANULAR, AGX_ERR_ARG, "running_cost: bad penalty_mode %d", penalty_mode);
  AGX_REQUIRE(error_mode == AGX_ERROR_BOX || M > 0, AGX_ERR_ARG, "running_cost: chamfer needs M > 0 target points");
  const size_t need = agx_running_cost_workspace_bytes(bsz, L);
  AGX_REQUIRE(workspace && workspace_bytes >= need, AGX_ERR_CAPACITY, "running_cost: workspace %zu < %zu bytes", workspace_bytes, need);
  const size_t smem = (size_t)(n + (error_mode == AGX_ERROR_CHAMFER ? M : 0)) * 12;
  AGX_REQUIRE(smem <= 200 * 1024, AGX_ERR_ARG, "running_cost: n + M = %d exceeds the shared-memory staging limit (17066 points)",
              n + (error_mode == AGX_ERROR_CHAMFER ? M : 0));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static thread_local size_t smem_set = 0;
  static thread_local DeviceOnce once;
  if (once.need()) smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    AGX_CUDA_OK(cudaFuncSetAttribute(running_cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  char* ws = static_cast<char*>(workspace);
  RewardArgs a{state, action, state_cur, bbox, target, bsz, L, n, action_dim, M, error_mode, penalty_mode, sim_real_ratio,
               reinterpret_cast<float*>(ws), reinterpret_cast<int*>(ws + align_up((size_t)bsz * L * 16, 256)), reward};
  { ProfScope ps(AGX_KIND_OTHER, st);
    running_cost_kernel<<<bsz * L, RW_THREADS, smem, st>>>(a); }
  AGX_LAUNCH_CHECK();
  return AGX_OK;
}

}  // extern "C"
