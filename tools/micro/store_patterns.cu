// Micro-benchmark: how fast can 148 CTAs x 16 warps stream [rows][160] fp32 tiles to / from HBM with the access patterns the
// tensor-core epilogues can produce?  (thread = row with 32-byte pieces, quad = row with 64-byte pieces, fully coalesced.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_patterns store_patterns.cu && ./store_patterns
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int FP = 160, TILE = 128, THREADS = 512;

__device__ __forceinline__ void stg256(float* p, const float (&v)[8]) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]),
               "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
__device__ __forceinline__ void ldg256(const float* p, float (&v)[8]) {
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]),
               "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p) : "memory");
}

// mode 0: thread = row, 16 columns per (chunk, half) as 2 x 256-bit   (the current epilogue)
// mode 1: quad of lanes = 64-byte row piece, 128-bit per lane          (after a 4x4 shuffle transpose)
// mode 2: warp = 128 contiguous floats of one row                      (fully coalesced, needs a shared-memory transpose)
// mode 3: thread = row, 32 columns per chunk as 4 x 256-bit (full 128-byte line per thread)
template <int MODE, bool STORE>
__global__ void __launch_bounds__(THREADS) pattern_kernel(float* buf, int n_tiles, float* sink) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = warp >> 3, q = warp & 3, half = (warp >> 2) & 1;
  float acc = 0.f;
  for (int t = blockIdx.x * 2 + slot; t < n_tiles; t += gridDim.x * 2) {
    float* tile = buf + (size_t)t * TILE * FP;
    if (MODE == 0) {
      float* row = tile + (size_t)(q * 32 + lane) * FP;
#pragma unroll
      for (int c = 0; c < 5; ++c)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float v[8];
          if (STORE) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = (float)(c + i);
            stg256(row + 32 * c + 16 * half + 8 * h, v);
          } else {
            ldg256(row + 32 * c + 16 * half + 8 * h, v);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc += v[i];
          }
        }
    } else if (MODE == 1) {
#pragma unroll
      for (int c = 0; c < 5; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float4* p = reinterpret_cast<float4*>(tile + (size_t)(q * 32 + j * 8 + (lane >> 2)) * FP + 32 * c + 16 * half + 4 * (lane & 3));
          if (STORE) *p = make_float4(c, j, lane, 1.f);
          else { const float4 v = *p; acc += v.x + v.y + v.z + v.w; }
        }
    } else if (MODE == 2) {
      // 16 warps of the CTA... this slot's 8 warps cover 128 rows x 160 floats = 5120 float4 -> 20 per lane per warp
#pragma unroll 4
      for (int i = 0; i < 20; ++i) {
        float4* p = reinterpret_cast<float4*>(tile) + (size_t)((warp & 7) * 20 + i) * 32 + lane;
        if (STORE) *p = make_float4(i, lane, 0.f, 1.f);
        else { const float4 v = *p; acc += v.x + v.y + v.z + v.w; }
      }
    } else if (MODE == 4 || MODE == 5) {
      // blocked layout [tile][piece][row][W]: W = 16 floats (mode 4: two 256-bit halves) or 8 floats (mode 5)
#pragma unroll
      for (int c = 0; c < 5; ++c)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float* p = MODE == 4 ? tile + (size_t)(c * 2 + half) * (TILE * 16) + (size_t)(q * 32 + lane) * 16 + 8 * h
                               : tile + (size_t)((c * 2 + half) * 2 + h) * (TILE * 8) + (size_t)(q * 32 + lane) * 8;
          float v[8];
          if (STORE) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = (float)(c + i);
            stg256(p, v);
          } else {
            ldg256(p, v);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc += v[i];
          }
        }
    } else {
      float* row = tile + (size_t)(q * 32 + lane) * FP;
      // half 0: chunks 0, 2, 4 ; half 1: chunks 1, 3
      for (int c = half; c < 5; c += 2)
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          float v[8];
          if (STORE) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = (float)(c + i);
            stg256(row + 32 * c + 8 * h, v);
          } else {
            ldg256(row + 32 * c + 8 * h, v);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc += v[i];
          }
        }
    }
  }
  if (acc == 123.456f) *sink = acc;
}

static int g_grid = 148;
template <int MODE, bool STORE>
void run(const char* name, float* buf, int n_tiles, float* sink) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 2; ++i) pattern_kernel<MODE, STORE><<<g_grid, THREADS>>>(buf, n_tiles, sink);
  cudaEventRecord(e0);
  const int reps = 5;
  for (int i = 0; i < reps; ++i) pattern_kernel<MODE, STORE><<<g_grid, THREADS>>>(buf, n_tiles, sink);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  ms /= reps;
  const double gb = (double)n_tiles * TILE * FP * 4 / 1e9;
  printf("grid %3d %-34s %s %7.3f ms  %7.1f GB/s  %5.1f B/clk/SM (%s)\n", g_grid, name, STORE ? "store" : "load ", ms, gb / (ms * 1e-3),
         gb / (ms * 1e-3) / g_grid / 1.965, cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char** argv) {
  if (argc > 1) g_grid = atoi(argv[1]);
  const int n_tiles = g_grid < 148 ? 14000 * g_grid / 148 : 14000;   // 1.15 GB, the relation encoder's C
  float *buf, *sink;
  cudaMalloc(&buf, (size_t)n_tiles * TILE * FP * 4);
  cudaMalloc(&sink, 4);
  cudaMemset(buf, 0, (size_t)n_tiles * TILE * FP * 4);
  run<0, true>("thread=row 2x256b per half-chunk", buf, n_tiles, sink);
  run<1, true>("quad=64B piece 128b", buf, n_tiles, sink);
  run<2, true>("coalesced 512B per warp instr", buf, n_tiles, sink);
  run<3, true>("thread=row full 128B line", buf, n_tiles, sink);
  run<4, true>("blocked [piece16][row][16]", buf, n_tiles, sink);
  run<5, true>("blocked [piece8][row][8]", buf, n_tiles, sink);
  run<4, false>("blocked [piece16][row][16]", buf, n_tiles, sink);
  run<5, false>("blocked [piece8][row][8]", buf, n_tiles, sink);
  run<0, false>("thread=row 2x256b per half-chunk", buf, n_tiles, sink);
  run<1, false>("quad=64B piece 128b", buf, n_tiles, sink);
  run<2, false>("coalesced 512B per warp instr", buf, n_tiles, sink);
  run<3, false>("thread=row full 128B line", buf, n_tiles, sink);
  return 0;
}
