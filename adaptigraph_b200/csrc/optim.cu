// Fused Adam over ONE flat fp32 bucket holding every parameter of the model.
// Replaces torch.optim.Adam(model.parameters(), lr=0.001) of dynamics/train/train.py:63 — its 22 per-tensor update chains
// become one launch over 252,903 floats; the same flat gradient bucket is what the data-parallel all-reduce acts on.
// Arithmetic = torch's single-tensor Adam (no amsgrad, no weight decay):
//   m <- m + (1 - b1) (g - m);  v <- b2 v + (1 - b2) g g;  p <- p - (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// The step count t lives in device memory so that a captured CUDA graph of the whole training step can be replayed.
#include "common.cuh"

namespace agx {

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, int64_t n, double lr_d, double b1_d, double b2_d, double eps_d,
                                                   float grad_scale, const int32_t* __restrict__ step) {
  // scalars as torch's Python-side Adam derives them: in double, rounded to fp32 where they meet the tensors
  const int t = *step + 1;
  const float step_size = (float)(lr_d / (1.0 - pow(b1_d, (double)t))), sqrt_bc2 = (float)sqrt(1.0 - pow(b2_d, (double)t));
  const float omb1 = (float)(1.0 - b1_d), b2 = (float)b2_d, omb2 = (float)(1.0 - b2_d), eps = (float)eps_d;
  const int64_t i0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
  if (i0 + 3 < n) {   // buckets are 16-byte aligned (torch allocations are 512-byte aligned)
    const float4 g4 = *reinterpret_cast<const float4*>(g + i0);
    float4 m4 = *reinterpret_cast<const float4*>(m + i0), v4 = *reinterpret_cast<const float4*>(v + i0), p4 = *reinterpret_cast<const float4*>(p + i0);
    float* gp = const_cast<float*>(reinterpret_cast<const float*>(&g4));
    float *mp = reinterpret_cast<float*>(&m4), *vp = reinterpret_cast<float*>(&v4), *pp = reinterpret_cast<float*>(&p4);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = gp[k] * grad_scale;
      mp[k] = mp[k] + omb1 * (gk - mp[k]);
      vp[k] = b2 * vp[k] + omb2 * gk * gk;
      pp[k] = pp[k] - step_size * (mp[k] / (sqrtf(vp[k]) / sqrt_bc2 + eps));
    }
    *reinterpret_cast<float4*>(m + i0) = m4;
    *reinterpret_cast<float4*>(v + i0) = v4;
    *reinterpret_cast<float4*>(p + i0) = p4;
  } else {
    for (int64_t i = i0; i < n; ++i) {
      const float gk = g[i] * grad_scale;
      const float mk = m[i] + omb1 * (gk - m[i]);
      const float vk = b2 * v[i] + omb2 * gk * gk;
      m[i] = mk; v[i] = vk;
      p[i] = p[i] - step_size * (mk / (sqrtf(vk) / sqrt_bc2 + eps));
    }
  }
}

__global__ void adam_tick_kernel(int32_t* step) { *step += 1; }

}  // namespace agx

extern "C" int agx_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1,
                             double beta2, double eps, float grad_scale, int32_t* step, agx_stream_t stream) {
  using namespace agx;
  AGX_REQUIRE(params && grads && exp_avg && exp_avg_sq && step, AGX_ERR_ARG, "adam_step: null pointer argument");
  AGX_REQUIRE(n > 0, AGX_ERR_ARG, "adam_step: n=%lld must be positive", (long long)n);
  AGX_REQUIRE(((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16 == 0, AGX_ERR_ARG,
              "adam_step: buckets must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t blocks = (n + 1023) / 1024;
  { ProfScope ps(AGX_KIND_OTHER, st);
    adam_kernel<<<(unsigned)blocks, 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, grad_scale, step); }
  AGX_LAUNCH_CHECK();
  { ProfScope ps(AGX_KIND_OTHER, st);
    adam_tick_kernel<<<1, 1, 0, st>>>(step); }
  AGX_LAUNCH_CHECK();
  return AGX_OK;
}
