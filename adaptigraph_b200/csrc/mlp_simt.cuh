// Exact-fp32 tile GEMM used by the fused MLP chains (AGX_PREC_FP32).
//
// One CTA of 256 threads owns a tile of TM=128 rows (edges or nodes) whose activations live
// in shared memory as Xs[row][k] (row stride LDX floats).  A layer is
//     acc[128][160] = Xs[128][K] * Wt[K][160]
// with Wt the k-major packed weight (common.cuh PackedLayout) streamed from L2 through a
// two-stage cp.async ring in chunks of KC rows.  Each thread owns an 8 x 10 register block:
// rows 8*tr..8*tr+7 (tr = tid/16) and columns {4tc..4tc+3, 64+4tc..64+4tc+3, 128+2tc, 128+2tc+1}
// (tc = tid%16), so every weight read is a conflict-free 128/64-bit shared load and every
// activation read is a warp-broadcast.  Epilogues run on the register block and either write the
// next layer's activations back into Xs in place or store rows to HBM.
#pragma once
#include "common.cuh"

namespace agx {

constexpr int TM = 128;            // rows per tile
constexpr int LDX = FP + 4;        // 164: row stride of Xs; (4*row + col) % 32 spreads float4 stores over banks
constexpr int MLP_THREADS = 256;
constexpr int KC = 8;              // weight rows per pipeline stage
constexpr int WS_STAGE = KC * FP;  // floats per stage
constexpr size_t MLP_SMEM_BYTES = (size_t)(TM * LDX + 2 * WS_STAGE) * sizeof(float);

struct Acc {
  float v[8][10];
};

__device__ __forceinline__ int acc_col(int tc, int c) {
  return c < 4 ? 4 * tc + c : (c < 8 ? 64 + 4 * tc + (c - 4) : 128 + 2 * tc + (c - 8));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

__device__ __forceinline__ void stage_weights(float* dst, const float* __restrict__ src, int tid) {
  // KC*FP floats = 320 float4
  for (int i = tid; i < WS_STAGE / 4; i += MLP_THREADS) cp_async16(dst + 4 * i, src + 4 * i);
  cp_async_commit();
}

// acc = Xs[:, 0:K] * Wt[0:K, :]   (K a multiple of KC).  Begins with a block barrier, so writes to
// Xs made before the call are visible and the weight ring is free.
template <int K>
__device__ __forceinline__ void tile_gemm(const float* Xs, const float* __restrict__ Wt, float* Ws, Acc& acc, int tid) {
  static_assert(K % KC == 0, "K must be a multiple of KC");
  const int tr = tid >> 4, tc = tid & 15;
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < 10; ++c) acc.v[r][c] = 0.f;
  __syncthreads();
  stage_weights(Ws, Wt, tid);
  constexpr int NC = K / KC;
  const float* xrow = Xs + (size_t)(8 * tr) * LDX;
#pragma unroll 1
  for (int ch = 0; ch < NC; ++ch) {
    cp_async_wait_all();
    __syncthreads();
    if (ch + 1 < NC) stage_weights(Ws + ((ch + 1) & 1) * WS_STAGE, Wt + (size_t)(ch + 1) * WS_STAGE, tid);
    const float* w = Ws + (ch & 1) * WS_STAGE;
    const float* x = xrow + ch * KC;
#pragma unroll
    for (int kk = 0; kk < KC; kk += 2) {
      float2 a[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) a[r] = *reinterpret_cast<const float2*>(x + r * LDX + kk);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float* wr = w + (kk + j) * FP;
        const float4 w0 = *reinterpret_cast<const float4*>(wr + 4 * tc);
        const float4 w1 = *reinterpret_cast<const float4*>(wr + 64 + 4 * tc);
        const float2 w2 = *reinterpret_cast<const float2*>(wr + 128 + 2 * tc);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const float av = j ? a[r].y : a[r].x;
          acc.v[r][0] = fmaf(av, w0.x, acc.v[r][0]);
          acc.v[r][1] = fmaf(av, w0.y, acc.v[r][1]);
          acc.v[r][2] = fmaf(av, w0.z, acc.v[r][2]);
          acc.v[r][3] = fmaf(av, w0.w, acc.v[r][3]);
          acc.v[r][4] = fmaf(av, w1.x, acc.v[r][4]);
          acc.v[r][5] = fmaf(av, w1.y, acc.v[r][5]);
          acc.v[r][6] = fmaf(av, w1.z, acc.v[r][6]);
          acc.v[r][7] = fmaf(av, w1.w, acc.v[r][7]);
          acc.v[r][8] = fmaf(av, w2.x, acc.v[r][8]);
          acc.v[r][9] = fmaf(av, w2.y, acc.v[r][9]);
        }
      }
    }
  }
}

// Loads the thread's 10 bias values.
__device__ __forceinline__ void load_bias(const float* __restrict__ bias, int tc, float (&b)[10]) {
  const float4 b0 = *reinterpret_cast<const float4*>(bias + 4 * tc);
  const float4 b1 = *reinterpret_cast<const float4*>(bias + 64 + 4 * tc);
  const float2 b2 = *reinterpret_cast<const float2*>(bias + 128 + 2 * tc);
  b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
  b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
  b[8] = b2.x; b[9] = b2.y;
}

// acc <- relu(acc + bias), written back into Xs in place (after a barrier: every thread has finished
// reading Xs for this layer).
__device__ __forceinline__ void epilogue_bias_relu_to_smem(float* Xs, Acc& acc, const float* __restrict__ bias, int tid) {
  const int tr = tid >> 4, tc = tid & 15;
  float b[10];
  load_bias(bias, tc, b);
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    float* row = Xs + (size_t)(8 * tr + r) * LDX;
    float o[10];
#pragma unroll
    for (int c = 0; c < 10; ++c) o[c] = fmaxf(acc.v[r][c] + b[c], 0.f);
    *reinterpret_cast<float4*>(row + 4 * tc) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(row + 64 + 4 * tc) = make_float4(o[4], o[5], o[6], o[7]);
    *reinterpret_cast<float2*>(row + 128 + 2 * tc) = make_float2(o[8], o[9]);
  }
}

// Stores the thread's register block rows to a row-major [rows][FP] HBM buffer.
__device__ __forceinline__ void store_rows(float* __restrict__ out, int64_t row0, int64_t n_rows, const Acc& acc, int tid) {
  const int tr = tid >> 4, tc = tid & 15;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int64_t gr = row0 + 8 * tr + r;
    if (gr < n_rows) {
      float* row = out + gr * FP;
      *reinterpret_cast<float4*>(row + 4 * tc) = make_float4(acc.v[r][0], acc.v[r][1], acc.v[r][2], acc.v[r][3]);
      *reinterpret_cast<float4*>(row + 64 + 4 * tc) = make_float4(acc.v[r][4], acc.v[r][5], acc.v[r][6], acc.v[r][7]);
      *reinterpret_cast<float2*>(row + 128 + 2 * tc) = make_float2(acc.v[r][8], acc.v[r][9]);
    }
  }
}

__device__ __forceinline__ void add_bias(Acc& acc, const float* __restrict__ bias, int tid) {
  float b[10];
  load_bias(bias, tid & 15, b);
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < 10; ++c) acc.v[r][c] += b[c];
}

}  // namespace agx
