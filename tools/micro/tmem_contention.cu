// Micro-benchmark: do tcgen05.mma (one issuing thread, M = 128, N = 96, K = 16, A in tensor memory) and the epilogue warps'
// tcgen05.ld / tcgen05.st share a resource?  One CTA per SM: warp 16 issues MMAs back to back (or not at all), E epilogue warps
// loop over  ld x16 x3 + wait  (mode 1),  st x8 x3 + wait  (mode 2)  or both (mode 3)  on accumulator-sized column ranges of their
// own lane quarter.  Reports cycles per MMA and tensor-memory bytes per cycle per SM moved by the epilogue warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I adaptigraph_b200/csrc -o tools/micro/tmem_contention tools/micro/tmem_contention.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace agx::tc;

__global__ void __launch_bounds__(544, 1) contention_kernel(long long* out, int n_epi, int mode, int mma_on, int iters) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tptr;
  __shared__ volatile int stop;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  for (int i = threadIdx.x; i < 160 * 160 * 2 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); stop = 0; }
  if (threadIdx.x < 32) tmem_alloc(&tptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 16) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_f16(128, 96);
      const uint64_t b0 = make_b_desc(smem_u32(smem), 128, (160 / 8) * 128);
      uint32_t parity = 0;
      const long long t0 = clock64();
      long long n = 0;
      if (mma_on) {
        for (int r = 0; r < iters; ++r) {
          for (int i = 0; i < 200; ++i) { const int ks = i % 10; mma_f16_ts(tmem, tmem + 96 + 8 * ks, b0 + 16ull * ks, idesc, ks > 0); }
          mma_commit(&bar);
          mbar_wait(&bar, parity);
          parity ^= 1;
          n += 200;
        }
      } else {
        while (clock64() - t0 < 200000ll * iters / 10) { }
      }
      const long long t1 = clock64();
      stop = 1;
      if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = n; }
    }
  } else if (warp < n_epi) {
    // lane quarter warp % 4; columns 256.. (away from the MMA's D [0, 96) and A [96, 176))
    const uint32_t t = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256 + 48 * (warp >> 2);
    uint32_t r[3][16];
    uint32_t acc = 0;
    long long bytes = 0;
    const long long t0 = clock64();
    while (!stop) {
      if (mode & 1) {
#pragma unroll
        for (int c = 0; c < 3; ++c) tmem_ld16(t + 16 * c, r[c]);
        tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < 3; ++c) acc += r[c][0] ^ r[c][15];
        bytes += 3 * 16 * 4 * 32;
      }
      if (mode & 2) {
        uint32_t v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = acc + i;
#pragma unroll
        for (int c = 0; c < 3; ++c) tmem_st8(t + 8 * c, v);
        tmem_wait_st();
        bytes += 3 * 8 * 4 * 32;
      }
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0 && lane == 0) { out[2 + 2 * warp] = bytes; out[3 + 2 * warp] = t1 - t0 + (acc == 0x12345u); }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* out;
  cudaMalloc(&out, 64 * 8);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t smem = 160 * 160 * 2 + 256;
  cudaFuncSetAttribute(contention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  printf("# epilogue warps, mode (1 ld, 2 st, 3 both), MMAs on -> cycles per MMA (floor 48), epilogue tensor-memory bytes per cycle per SM\n");
  for (int mma_on = 0; mma_on <= 1; ++mma_on)
    for (int mode = 1; mode <= 3; ++mode)
      for (int n_epi : {0, 4, 8, 16}) {
        if (n_epi == 0 && (mode != 1 || !mma_on)) continue;
        cudaMemset(out, 0, 64 * 8);
        contention_kernel<<<sms, 544, smem>>>(out, n_epi, mode, mma_on, 20);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
        long long h[64];
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        double bpc = 0;
        for (int w = 0; w < n_epi; ++w) if (h[3 + 2 * w] > 0) bpc += (double)h[2 + 2 * w] / h[3 + 2 * w];
        printf("epi_warps %2d  mode %d  mma %d  cycles/MMA %6.1f  epilogue B/cycle %7.1f\n", n_epi, mode, mma_on, h[1] ? (double)h[0] / h[1] : 0.0, bpc);
      }
  return 0;
}
