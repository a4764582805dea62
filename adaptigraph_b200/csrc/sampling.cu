// Farthest-point sampling on the device: the step that produces the model's particles from raw point
// clouds / simulator states.  Replaces, with identical selections,
//
//   * dgl.geometry.farthest_point_sampler(pos, npoints, start_idx) as called at dynamics/dataset/graph.py:11-12 and
//     planning/perception.py:271 (third-party; restated from DGL's CPU operator, see oracle/sampling_oracle.py):
//     squared distances ((dx*dx + dy*dy) + dz*dz, fp32, no FMA contraction), d[j] = min over the chosen points, next = the
//     FIRST index of the maximum;
//   * fps_rad_idx(pcd, radius) of dynamics/utils.py:10-24: the same recurrence on np.linalg.norm distances (fp32 sqrt of the
//     same sum), continued while max_j d[j] > radius (compared in double, as numpy compares a float32 with a Python float).
//
// One CTA per cloud: the points and their running distances stay in shared memory; one iteration = one pass over the points
// (each thread its strided share), a warp-shuffle arg-max, a 32-entry shared-memory exchange and one more warp arg-max.
#include <cooperative_groups.h>

#include "common.cuh"

namespace agx {

constexpr int FPS_THREADS = 1024;
constexpr unsigned FPS_FULL = 0xffffffffu;

// (value, index) arg-max with ties towards the lower index
__device__ __forceinline__ void argmax_merge(float& v, int& i, float ov, int oi) {
  if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}

template <bool SQRT_DOMAIN>
__global__ void __launch_bounds__(FPS_THREADS) fps_kernel(const float* __restrict__ pos, const int32_t* __restrict__ n_points, int N,
                                                          int max_samples, const int32_t* __restrict__ start_idx, double radius,
                                                          const double* __restrict__ radii /* per cloud, or null */,
                                                          int32_t* __restrict__ idx_out, int32_t* __restrict__ n_out) {
  extern __shared__ float fps_smem[];
  float* px = fps_smem;
  float* py = px + N;
  float* pz = py + N;
  float* dist = pz + N;
  __shared__ float red_v[32];
  __shared__ int red_i[32];
  __shared__ int cur_s;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (radii) radius = radii[b];
  const int n = n_points ? min(max(n_points[b], 0), N) : N;
  const float* p = pos + (size_t)b * N * 3;
  int32_t* out = idx_out + (size_t)b * max_samples;
  if (n == 0 || max_samples == 0) { if (tid == 0) n_out[b] = 0; return; }
  const float INF = __int_as_float(0x7f800000);
  for (int j = tid; j < n; j += FPS_THREADS) { px[j] = p[3 * j]; py[j] = p[3 * j + 1]; pz[j] = p[3 * j + 2]; dist[j] = INF; }
  if (tid == 0) {
    const int s = min(max(start_idx[b], 0), n - 1);
    cur_s = s;
    out[0] = s;
  }
  __syncthreads();
  int count = 1;
  const int limit = min(max_samples, SQRT_DOMAIN ? n : max_samples);
  while (true) {
    const int cur = cur_s;
    const float cx = px[cur], cy = py[cur], cz = pz[cur];
    float bv = -1.f;
    int bi = 0x7fffffff;
    for (int j = tid; j < n; j += FPS_THREADS) {
      const float dx = __fsub_rn(px[j], cx), dy = __fsub_rn(py[j], cy), dz = __fsub_rn(pz[j], cz);
      float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      if (SQRT_DOMAIN) d = __fsqrt_rn(d);
      const float m = fminf(dist[j], d);
      dist[j] = m;
      if (m > bv) { bv = m; bi = j; }   // ascending j: strict > keeps the first maximum
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) argmax_merge(bv, bi, __shfl_xor_sync(FPS_FULL, bv, o), __shfl_xor_sync(FPS_FULL, bi, o));
    if (lane == 0) { red_v[warp] = bv; red_i[warp] = bi; }
    __syncthreads();
    bv = red_v[lane];
    bi = red_i[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) argmax_merge(bv, bi, __shfl_xor_sync(FPS_FULL, bv, o), __shfl_xor_sync(FPS_FULL, bi, o));
    // every warp now holds the same (bv, bi)
    bool stop = count >= limit;
    if (SQRT_DOMAIN) stop = stop || !((double)bv > radius);   // utils.py:17  while dist.max() > radius
    if (stop) break;
    if (tid == 0) { cur_s = bi; out[count] = bi; }
    ++count;
    __syncthreads();
  }
  if (tid == 0) n_out[b] = count;
}

// Clouds that do not fit one CTA's shared memory: a thread-block CLUSTER per cloud.  CTA q of the cluster keeps points
// [q * chunk, (q + 1) * chunk) and their running distances in its own shared memory; per pick every CTA scans its share, publishes its
// local (max, arg-max) in a double-buffered slot, the cluster synchronises once, and every CTA reads all slots — and the winner's
// coordinates — through distributed shared memory.  Same arithmetic and tie rule as fps_kernel, so the picks are identical.
template <bool SQRT_DOMAIN>
__global__ void __launch_bounds__(FPS_THREADS) fps_cluster_kernel(const float* __restrict__ pos, const int32_t* __restrict__ n_points, int N,
                                                                  int chunk, int max_samples, const int32_t* __restrict__ start_idx,
                                                                  double radius, const double* __restrict__ radii,
                                                                  int32_t* __restrict__ idx_out, int32_t* __restrict__ n_out) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ float fps_smem[];
  float* px = fps_smem;
  float* py = px + chunk;
  float* pz = py + chunk;
  float* dist = pz + chunk;
  __shared__ float red_v[32];
  __shared__ int red_i[32];
  __shared__ float best_v[2];
  __shared__ int best_i[2];
  const int CL = (int)cluster.num_blocks(), q = (int)cluster.block_rank();
  const int b = blockIdx.x / CL, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (radii) radius = radii[b];
  const int n = n_points ? min(max(n_points[b], 0), N) : N;
  const float* p = pos + (size_t)b * N * 3;
  int32_t* out = idx_out + (size_t)b * max_samples;
  if (n == 0 || max_samples == 0) { if (tid == 0 && q == 0) n_out[b] = 0; return; }   // uniform over the cluster: nobody syncs
  const int lo = q * chunk, cnt_local = max(0, min(n - lo, chunk));
  const float INF = __int_as_float(0x7f800000);
  for (int j = tid; j < cnt_local; j += FPS_THREADS) {
    px[j] = p[3 * (lo + j)]; py[j] = p[3 * (lo + j) + 1]; pz[j] = p[3 * (lo + j) + 2]; dist[j] = INF;
  }
  int cur = min(max(start_idx[b], 0), n - 1);
  if (tid == 0 && q == 0) out[0] = cur;
  cluster.sync();                                  // every CTA's points are staged before anyone reads them remotely
  int count = 1;
  const int limit = min(max_samples, SQRT_DOMAIN ? n : max_samples);
  for (int it = 0;; ++it) {
    const int owner = cur / chunk, off = cur - owner * chunk;
    const float cx = cluster.map_shared_rank(px, owner)[off], cy = cluster.map_shared_rank(py, owner)[off],
                cz = cluster.map_shared_rank(pz, owner)[off];
    float bv = -1.f;
    int bi = 0x7fffffff;
    for (int j = tid; j < cnt_local; j += FPS_THREADS) {
      const float dx = __fsub_rn(px[j], cx), dy = __fsub_rn(py[j], cy), dz = __fsub_rn(pz[j], cz);
      float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      if (SQRT_DOMAIN) d = __fsqrt_rn(d);
      const float m = fminf(dist[j], d);
      dist[j] = m;
      if (m > bv) { bv = m; bi = lo + j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) argmax_merge(bv, bi, __shfl_xor_sync(FPS_FULL, bv, o), __shfl_xor_sync(FPS_FULL, bi, o));
    if (lane == 0) { red_v[warp] = bv; red_i[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
      bv = red_v[lane];
      bi = red_i[lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) argmax_merge(bv, bi, __shfl_xor_sync(FPS_FULL, bv, o), __shfl_xor_sync(FPS_FULL, bi, o));
      if (lane == 0) { best_v[it & 1] = bv; best_i[it & 1] = bi; }
    }
    cluster.sync();                                // all local results of this pick are published (and red_* may be reused)
    bv = -1.f;
    bi = 0x7fffffff;
    for (int c = 0; c < CL; ++c)                   // same order in every thread of every CTA: identical winner everywhere
      argmax_merge(bv, bi, cluster.map_shared_rank(best_v, c)[it & 1], cluster.map_shared_rank(best_i, c)[it & 1]);
    bool stop = count >= limit;
    if (SQRT_DOMAIN) stop = stop || !((double)bv > radius);
    if (stop) break;
    if (tid == 0 && q == 0) out[count] = bi;
    cur = bi;
    ++count;
  }
  if (tid == 0 && q == 0) n_out[b] = count;
  cluster.sync();                                  // nobody leaves while a peer may still read its shared memory
}

constexpr size_t FPS_CTA_SMEM = 200 * 1024;        // 12800 points per CTA
constexpr int FPS_MAX_CLUSTER = 16;               // non-portable cluster size: 204800 points per cloud

static int fps_cluster_launch(const float* pos, const int32_t* n_points, int B, int N, int max_samples, const int32_t* start_idx,
                              double radius, const double* radii, int32_t* idx_out, int32_t* n_out, cudaStream_t st) {
  int CL = 2;
  while (CL < FPS_MAX_CLUSTER && (size_t)((N + CL - 1) / CL) * 16 > FPS_CTA_SMEM) CL *= 2;
  const int chunk = (N + CL - 1) / CL;
  const size_t smem = (size_t)chunk * 16;
  AGX_REQUIRE(smem <= FPS_CTA_SMEM, AGX_ERR_ARG, "fps: N=%d exceeds the cluster staging limit (%d points per cloud)", N,
              (int)(FPS_MAX_CLUSTER * (FPS_CTA_SMEM / 16)));
  auto launch = [&](auto kernel) -> int {
    AGX_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (CL > 8) AGX_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * CL));
    cfg.blockDim = dim3(FPS_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    ProfScope ps(AGX_KIND_OTHER, st);
    AGX_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, pos, n_points, N, chunk, max_samples, start_idx, radius, radii, idx_out, n_out));
    return AGX_OK;
  };
  if (int rc = radius < 0.0 ? launch(fps_cluster_kernel<false>) : launch(fps_cluster_kernel<true>)) return rc;
  AGX_LAUNCH_CHECK();
  return AGX_OK;
}

}  // namespace agx

extern "C" {

static int fps_impl(const float* pos, const int32_t* n_points, int32_t B, int32_t N, int32_t max_samples, const int32_t* start_idx,
                    double radius, const double* radii, int32_t* idx_out, int32_t* n_out, agx_stream_t stream) {
  using namespace agx;
  AGX_REQUIRE(pos && start_idx && idx_out && n_out, AGX_ERR_ARG, "fps: null pointer argument");
  AGX_REQUIRE(B > 0 && N > 0 && max_samples > 0, AGX_ERR_ARG, "fps: B=%d N=%d max_samples=%d must be positive", B, N, max_samples);
  cudaStream_t st0 = static_cast<cudaStream_t>(stream);
  if ((size_t)N * 16 > FPS_CTA_SMEM) return agx::fps_cluster_launch(pos, n_points, B, N, max_samples, start_idx, radius, radii, idx_out, n_out, st0);
  const size_t smem = (size_t)N * 16;
  cudaStream_t st = st0;
  static thread_local size_t set_count = 0, set_radius = 0;
  static thread_local DeviceOnce once;
  if (once.need()) set_count = set_radius = 0;
  if (radius < 0.0) {
    if (smem > 48 * 1024 && smem > set_count) {
      AGX_CUDA_OK(cudaFuncSetAttribute(fps_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      set_count = smem;
    }
    ProfScope ps(AGX_KIND_OTHER, st);
    fps_kernel<false><<<B, FPS_THREADS, smem, st>>>(pos, n_points, N, max_samples, start_idx, radius, nullptr, idx_out, n_out);
  } else {
    if (smem > 48 * 1024 && smem > set_radius) {
      AGX_CUDA_OK(cudaFuncSetAttribute(fps_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      set_radius = smem;
    }
    ProfScope ps(AGX_KIND_OTHER, st);
    fps_kernel<true><<<B, FPS_THREADS, smem, st>>>(pos, n_points, N, max_samples, start_idx, radius, radii, idx_out, n_out);
  }
  AGX_LAUNCH_CHECK();
  return AGX_OK;
}

int agx_fps(const float* pos, const int32_t* n_points, int32_t B, int32_t N, int32_t max_samples, const int32_t* start_idx,
            double radius, int32_t* idx_out, int32_t* n_out, agx_stream_t stream) {
  return fps_impl(pos, n_points, B, N, max_samples, start_idx, radius, nullptr, idx_out, n_out, stream);
}

int agx_fps_radii(const float* pos, const int32_t* n_points, int32_t B, int32_t N, int32_t max_samples, const int32_t* start_idx,
                  const double* radii, int32_t* idx_out, int32_t* n_out, agx_stream_t stream) {
  AGX_REQUIRE(radii, AGX_ERR_ARG, "fps_radii: null radius array");
  return fps_impl(pos, n_points, B, N, max_samples, start_idx, 0.0, radii, idx_out, n_out, stream);
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ chamfer distance
// planning/losses.py:4-10 — the MPC error term evaluated on the rollout's last frame (plan.py:36, :146): for every sample b
//   mean_m min_n |x[b,n] - y[m]|  +  mean_n min_m |x[b,n] - y[m]|.
// The reference materialises two (B, M, N, 3) tensors; here one CTA per sample keeps both point sets in shared memory, a thread
// owns one point of one set and scans the other (min of squared distances, one sqrt per point: sqrt is monotone), then a fixed-order
// block reduction gives the two means.
namespace agx {

constexpr int CH_THREADS = 256;

__device__ __forceinline__ float block_sum_256(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < CH_THREADS / 32; ++w) t += red[w];   // same order in every thread
  return t;
}

__global__ void __launch_bounds__(CH_THREADS) chamfer_kernel(const float* __restrict__ x, const float* __restrict__ y, int N, int M,
                                                             int64_t y_stride_b, float* __restrict__ out) {
  extern __shared__ float ch_smem[];
  __shared__ float red[CH_THREADS / 32];
  float* xs = ch_smem;            // [N][3]
  float* ys = ch_smem + 3 * N;    // [M][3]
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* xb = x + (size_t)b * N * 3;
  const float* yb = y + (size_t)b * y_stride_b;
  for (int i = tid; i < 3 * N; i += CH_THREADS) xs[i] = xb[i];
  for (int i = tid; i < 3 * M; i += CH_THREADS) ys[i] = yb[i];
  __syncthreads();
  const float INF = __int_as_float(0x7f800000);
  float sum_m = 0.f, sum_n = 0.f;
  for (int m = tid; m < M; m += CH_THREADS) {
    const float a0 = ys[3 * m], a1 = ys[3 * m + 1], a2 = ys[3 * m + 2];
    float best = INF;
    for (int n = 0; n < N; ++n) {
      const float d0 = xs[3 * n] - a0, d1 = xs[3 * n + 1] - a1, d2 = xs[3 * n + 2] - a2;
      best = fminf(best, __fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)));
    }
    sum_m += __fsqrt_rn(best);
  }
  for (int n = tid; n < N; n += CH_THREADS) {
    const float a0 = xs[3 * n], a1 = xs[3 * n + 1], a2 = xs[3 * n + 2];
    float best = INF;
    for (int m = 0; m < M; ++m) {
      const float d0 = a0 - ys[3 * m], d1 = a1 - ys[3 * m + 1], d2 = a2 - ys[3 * m + 2];
      best = fminf(best, __fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)));
    }
    sum_n += __fsqrt_rn(best);
  }
  const float tm = block_sum_256(sum_m, red), tn = block_sum_256(sum_n, red);
  if (tid == 0) out[b] = tm / (float)M + tn / (float)N;
}

}  // namespace agx

extern "C" int agx_chamfer(const float* x, const float* y, int32_t B, int32_t N, int32_t M, int32_t y_batched, float* out,
                           agx_stream_t stream) {
  using namespace agx;
  AGX_REQUIRE(x && y && out, AGX_ERR_ARG, "chamfer: null pointer argument");
  AGX_REQUIRE(B > 0 && N > 0 && M > 0, AGX_ERR_ARG, "chamfer: B=%d N=%d M=%d must be positive", B, N, M);
  const size_t smem = (size_t)(N + M) * 12;
  AGX_REQUIRE(smem <= 200 * 1024, AGX_ERR_ARG, "chamfer: N + M = %d exceeds the shared-memory staging limit (17066 points)", N + M);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static thread_local size_t smem_set = 0;
  static thread_local DeviceOnce once;
  if (once.need()) smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    AGX_CUDA_OK(cudaFuncSetAttribute(chamfer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  { ProfScope ps(AGX_KIND_OTHER, st);
    chamfer_kernel<<<B, CH_THREADS, smem, st>>>(x, y, N, M, y_batched ? (int64_t)M * 3 : 0, out); }
  AGX_LAUNCH_CHECK();
  return AGX_OK;
}
