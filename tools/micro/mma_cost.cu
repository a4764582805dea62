// Micro-benchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, M = 128, K = 16, A from tensor memory, B from shared memory,
// no swizzle K-major: the chains' MMA) as a function of N, issued back to back by one thread -- and by two threads of two warps at
// once, as the two slots of a chain CTA do.  One CTA per SM on all SMs (so that clocks / power are those of a real kernel).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I adaptigraph_b200/csrc -o tools/micro/mma_cost tools/micro/mma_cost.cu && tools/micro/mma_cost
#include <cstdio>
#include <cstdint>
#include <string>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace agx::tc;

__global__ void __launch_bounds__(128, 1) mma_cost_kernel(long long* out, int N, int per_commit, int reps, int issuers) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ uint32_t tptr;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  for (int i = threadIdx.x; i < 256 * 160 * 2 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 ones
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tptr;
  const int warp = threadIdx.x >> 5;
  if (warp < issuers && elect_one()) {
    const uint32_t idesc = make_idesc_f16(128, N);
    const uint32_t sbo = (160 / 8) * 128;
    const uint64_t b0 = make_b_desc(smem_u32(smem), 128, sbo);
    const uint32_t d = tmem + (warp ? 256u : 0u), a0 = tmem + (warp ? 256u : 0u) + 160u;   // per issuer: D [0, 160), A [160, 240)
    uint32_t parity = 0;
    long long t0 = 0;
    for (int r = -2; r < reps; ++r) {
      if (r == 0) t0 = clock64();
      for (int i = 0; i < per_commit; ++i) {
        const int ks = i % 10;
        mma_f16_ts(d, a0 + 8 * ks, b0 + 16ull * ks, idesc, ks > 0);
      }
      mma_commit(&bar[warp]);
      mbar_wait(&bar[warp], parity);
      parity ^= 1;
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[warp] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* out;
  cudaMalloc(&out, 16);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t smem = 256 * 160 * 2 + 256;
  cudaFuncSetAttribute(mma_cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  printf("# cycles per MMA (M=128, K=16, kind::f16, TS): N, MMAs per commit, issuing threads -> cycles/MMA (per issuer), floor 128*N/256\n");
  const int Ns[] = {32, 64, 80, 96, 128, 160, 256};
  for (int issuers = 1; issuers <= 2; ++issuers)
    for (int pc : {20, 200})
      for (int N : Ns) {
        if (issuers == 2 && N > 160) continue;
        const int reps = 4000 / pc;
        cudaMemset(out, 0, 16);
        mma_cost_kernel<<<sms, 128, smem>>>(out, N, pc, reps, issuers);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("N=%d: %s\n", N, cudaGetErrorString(e)); return 1; }
        long long h[2];
        cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
        printf("N %3d  per_commit %3d  issuers %d  cycles/MMA %.1f %s  floor %.0f\n", N, pc, issuers, (double)h[0] / (reps * pc),
               issuers == 2 ? (std::to_string((double)h[1] / (reps * pc)).c_str()) : "", 128.0 * N / 256);
      }
  return 0;
}
