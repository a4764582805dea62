// Training path: DynamicsPredictor.forward with saved activations, and its backward
// (reference: dynamics/gnn/model.py:129-313 differentiated by torch autograd in dynamics/train/train.py:90-112).
//
// Exact fp32 FFMA tiles throughout (mlp_simt.cuh).  The forward is the same algebra as forward.cu (CSR relations,
// propagators split by operand) evaluated layer by layer so that every activation the backward needs is kept:
//
//   nodes : p_in -> h1 -> h2 -> penc ; A_n = W_enc*penc + b ; P_0 = penc
//   edges : rel_in -> g1 -> g2 -> renc ; C = W_rel*renc + b
//   pstep k: Qr_k = P_k*W_recv^T, Qs_k = P_k*W_send^T ; agg_k[n] = sum_e relu(C_e + Qr_k[n] + Qs_k[send e])
//            P_{k+1} = relu(A_n + agg_k*W_agg^T + P_k)
//   head  : u1, u2, motion ; pred_pos = pos + clamp(motion)
//
// Backward: dgrad GEMMs with the ReLU mask applied while the upstream gradient tile is staged, weight gradients as
// per-CTA partial 160x160 products reduced in a fixed order (deterministic, no atomics), relation-side gradients by
// segmented sums over the receiver CSR and a sender-sorted permutation of the same relations (CSC), and the gradient
// with respect to `state` (BPTT through the history features and pred_pos = state[:, -1] + ...).
#include "common.cuh"
#include "mlp_simt.cuh"
#include "tc_chain.cuh"
#include "tc_backward.cuh"
#include "tc_forward.cuh"
#include "tc_wgrad.cuh"

namespace agx {

constexpr float MOTION_CLAMP_T = 100.f;

// ------------------------------------------------------------------------------------ generic linear layer
// Y[m][0:160] (+)= [relu]( (X[m][0:K] (*) [mask[m][k] > 0]) * Wt[0:K][0:160] + bias + add1[m] + add2[m] )
struct LinArgs {
  const float* X; int ldx;
  const float* mask; int ldm;        // optional: X is multiplied elementwise by (mask > 0)   (ReLU backward)
  const float* Wt;                   // k-major [K][160]
  const float* bias;                 // optional [160]
  const float* add1; const float* add2;   // optional [M][160]
  float* Y; int ldy;
  int64_t M;
  int relu, accumulate;
  int n_store;                       // number of output columns stored (<= 160)
};

template <int K>
__global__ void __launch_bounds__(MLP_THREADS, 2) lin_kernel(const LinArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* Xs = smem;
  float* Ws = smem + TM * LDX;
  const int tid = threadIdx.x, tr = tid >> 4, tc = tid & 15;
  const int n_tiles = (int)((a.M + TM - 1) / TM);
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = (int64_t)tile * TM;
    __syncthreads();
    for (int i = tid; i < TM * (K / 4); i += MLP_THREADS) {
      const int row = i / (K / 4), c4 = i - row * (K / 4);
      const int64_t gr = row0 + row;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gr < a.M) {
        v = *reinterpret_cast<const float4*>(a.X + gr * a.ldx + 4 * c4);
        if (a.mask) {
          const float4 m = *reinterpret_cast<const float4*>(a.mask + gr * a.ldm + 4 * c4);
          v.x = m.x > 0.f ? v.x : 0.f; v.y = m.y > 0.f ? v.y : 0.f; v.z = m.z > 0.f ? v.z : 0.f; v.w = m.w > 0.f ? v.w : 0.f;
        }
      }
      *reinterpret_cast<float4*>(Xs + (size_t)row * LDX + 4 * c4) = v;
    }
    Acc acc;
    tile_gemm<K>(Xs, a.Wt, Ws, acc, tid);
    float b[10] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (a.bias) load_bias(a.bias, tc, b);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int64_t gr = row0 + 8 * tr + r;
      if (gr >= a.M) continue;
#pragma unroll
      for (int c = 0; c < 10; ++c) {
        const int col = acc_col(tc, c);
        if (col >= a.n_store) continue;
        float v = acc.v[r][c] + b[c];
        if (a.add1) v += a.add1[gr * FP + col];
        if (a.add2) v += a.add2[gr * FP + col];
        if (a.relu) v = fmaxf(v, 0.f);
        float* y = a.Y + gr * a.ldy + col;
        *y = a.accumulate ? *y + v : v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------ weight gradient partials
// part[cta][n][j] = sum over the CTA's rows m of (dY[m][n] * [mask[m][n] > 0]) * X[m][j],  part_b[cta][n] = sum_m dY*mask
struct WgradArgs {
  const float* dY; int ldd; const float* mask; int ldm; const float* X; int ldx; int kx;   // kx = valid X columns (multiple of 4, <= 160)
  int64_t M;
  float* part;      // [grid][160*160 + 160]
};
constexpr int WG_LD = FP + 4;
constexpr size_t WG_SMEM = (size_t)2 * TM * WG_LD * sizeof(float);

__global__ void __launch_bounds__(MLP_THREADS, 1) wgrad_kernel(const WgradArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* Ds = smem;                 // [128][164] upstream gradient (masked)
  float* Xs = smem + TM * WG_LD;    // [128][164] layer input
  const int tid = threadIdx.x, tn = tid >> 4, tj = tid & 15;
  float acc[10][10];
  float accb[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    accb[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 10; ++j) acc[i][j] = 0.f;
  }
  const int n_tiles = (int)((a.M + TM - 1) / TM);
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = (int64_t)tile * TM;
    __syncthreads();
    for (int i = tid; i < TM * (FP / 4); i += MLP_THREADS) {
      const int row = i / (FP / 4), c4 = i - row * (FP / 4);
      const int64_t gr = row0 + row;
      float4 d = make_float4(0.f, 0.f, 0.f, 0.f), x = d;
      if (gr < a.M) {
        d = *reinterpret_cast<const float4*>(a.dY + gr * a.ldd + 4 * c4);
        if (a.mask) {
          const float4 m = *reinterpret_cast<const float4*>(a.mask + gr * a.ldm + 4 * c4);
          d.x = m.x > 0.f ? d.x : 0.f; d.y = m.y > 0.f ? d.y : 0.f; d.z = m.z > 0.f ? d.z : 0.f; d.w = m.w > 0.f ? d.w : 0.f;
        }
        if (4 * c4 < a.kx) x = *reinterpret_cast<const float4*>(a.X + gr * a.ldx + 4 * c4);
      }
      *reinterpret_cast<float4*>(Ds + (size_t)row * WG_LD + 4 * c4) = d;
      *reinterpret_cast<float4*>(Xs + (size_t)row * WG_LD + 4 * c4) = x;
    }
    __syncthreads();
#pragma unroll 2
    for (int m = 0; m < TM; ++m) {
      const float* dr = Ds + (size_t)m * WG_LD;
      const float* xr = Xs + (size_t)m * WG_LD;
      const float4 d0 = *reinterpret_cast<const float4*>(dr + 4 * tn), d1 = *reinterpret_cast<const float4*>(dr + 64 + 4 * tn);
      const float2 d2 = *reinterpret_cast<const float2*>(dr + 128 + 2 * tn);
      const float4 x0 = *reinterpret_cast<const float4*>(xr + 4 * tj), x1 = *reinterpret_cast<const float4*>(xr + 64 + 4 * tj);
      const float2 x2 = *reinterpret_cast<const float2*>(xr + 128 + 2 * tj);
      const float dv[10] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w, d2.x, d2.y};
      const float xv[10] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w, x2.x, x2.y};
#pragma unroll
      for (int i = 0; i < 10; ++i) {
#pragma unroll
        for (int j = 0; j < 10; ++j) acc[i][j] = fmaf(dv[i], xv[j], acc[i][j]);
        if (tj == 0) accb[i] += dv[i];
      }
    }
  }
  float* out = a.part + (size_t)blockIdx.x * (FP * FP + FP);
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const int n = acc_col(tn, i);
#pragma unroll
    for (int j = 0; j < 10; ++j) out[n * FP + acc_col(tj, j)] = acc[i][j];
    if (tj == 0) out[FP * FP + n] = accb[i];
  }
}

// dW[n*ld + col0 + j] += sum_cta part[cta][n][j] (n < F, j < K);  db[n] += sum_cta part_b[cta][n]   (fixed order)
__global__ void wgrad_reduce_kernel(const float* __restrict__ part, int n_cta, int F, int K, int ld, int col0, float* __restrict__ dW,
                                    float* __restrict__ db) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < F * K) {
    const int n = i / K, j = i - n * K;
    float s = 0.f;
    for (int c = 0; c < n_cta; ++c) s += part[(size_t)c * (FP * FP + FP) + n * FP + j];
    dW[(size_t)n * ld + col0 + j] += s;
  } else if (db && i < F * K + F) {
    const int n = i - F * K;
    float s = 0.f;
    for (int c = 0; c < n_cta; ++c) s += part[(size_t)c * (FP * FP + FP) + FP * FP + n];
    db[n] += s;
  }
}

// ------------------------------------------------------------------------------------ feature assembly
// nfeat / p_in per node (model.py:155-195)
__global__ void node_prep_kernel(const float* __restrict__ state, const float* __restrict__ attrs, const float* __restrict__ action,
                                 const float* __restrict__ p_instance, const float* __restrict__ physics, int B, int N, int n_p,
                                 float* __restrict__ nfeat, float* __restrict__ p_in) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= (int64_t)B * N) return;
  const int b = (int)(r / N), n = (int)(r - (int64_t)b * N);
  float s[H_FIX][3];
#pragma unroll
  for (int h = 0; h < H_FIX; ++h) {
    const float* p = state + (((size_t)b * H_FIX + h) * N + n) * 3;
    s[h][0] = p[0]; s[h][1] = p[1]; s[h][2] = p[2];
  }
  const float a0 = attrs[r * 2], a1 = attrs[r * 2 + 1];
  const float grp = n < n_p ? p_instance[(size_t)b * n_p + n] : 0.f;
  float4* nf = reinterpret_cast<float4*>(nfeat + r * NFEAT);
  nf[0] = make_float4(s[1][0] - s[0][0], s[1][1] - s[0][1], s[1][2] - s[0][2], s[2][0] - s[1][0]);
  nf[1] = make_float4(s[2][1] - s[1][1], s[2][2] - s[1][2], s[3][0] - s[2][0], s[3][1] - s[2][1]);
  nf[2] = make_float4(s[3][2] - s[2][2], s[3][0], s[3][1], s[3][2]);
  nf[3] = make_float4(a0, a1, grp, 0.f);
  float4* pi = reinterpret_cast<float4*>(p_in + r * D_NODE_IN);
  pi[0] = make_float4(a0, a1, n < n_p ? physics[b] : 0.f, action[r * 3]);
  pi[1] = make_float4(action[r * 3 + 1], action[r * 3 + 2], 0.f, 0.f);
}

// rel_in per relation (model.py:224-253)
__global__ void edge_prep_kernel(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ send, const int32_t* __restrict__ recv,
                                 int64_t rows, int N, int64_t E_cap, const float* __restrict__ nfeat, float* __restrict__ rel_in) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t E = min((int64_t)row_ptr[rows], E_cap);
  if (e >= E_cap) return;
  float in[D_REL_IN];
#pragma unroll
  for (int i = 0; i < D_REL_IN; ++i) in[i] = 0.f;
  if (e < E) {
    const int r = recv[e];
    const int s = (r / N) * N + send[e];
    const float4* fr = reinterpret_cast<const float4*>(nfeat + (size_t)r * NFEAT);
    const float4* fs = reinterpret_cast<const float4*>(nfeat + (size_t)s * NFEAT);
    const float4 r0 = fr[0], r1 = fr[1], r2 = fr[2], r3 = fr[3];
    const float4 s0 = fs[0], s1 = fs[1], s2 = fs[2], s3 = fs[3];
    in[0] = r3.x; in[1] = r3.y; in[2] = s3.x; in[3] = s3.y;
    in[4] = fabsf(r3.z - s3.z);
    in[5] = r0.x - s0.x; in[6] = r0.y - s0.y; in[7] = r0.z - s0.z; in[8] = r0.w - s0.w;
    in[9] = r1.x - s1.x; in[10] = r1.y - s1.y; in[11] = r1.z - s1.z; in[12] = r1.w - s1.w;
    in[13] = r2.x - s2.x; in[14] = r2.y - s2.y; in[15] = r2.z - s2.z; in[16] = r2.w - s2.w;
  }
  float4* o = reinterpret_cast<float4*>(rel_in + e * D_REL_IN);
#pragma unroll
  for (int i = 0; i < D_REL_IN / 4; ++i) o[i] = make_float4(in[4 * i], in[4 * i + 1], in[4 * i + 2], in[4 * i + 3]);
}

// ------------------------------------------------------------------------------------ relation effects (forward / backward)
constexpr int TA_NODES = 8;
constexpr int TA_THREADS = TA_NODES * (FP / 4);

__device__ __forceinline__ float4 pre3(const float4 c, const float4 qr, const float4 qs) {
  return make_float4((c.x + qr.x) + qs.x, (c.y + qr.y) + qs.y, (c.z + qr.z) + qs.z, (c.w + qr.w) + qs.w);
}

// agg[n] = sum_{e in row n} relu(C_e + Qr[n] + Qs[send e])
__global__ void __launch_bounds__(TA_THREADS) train_aggregate_kernel(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ send,
                                                                    int64_t rows, int N, int64_t E_cap, const float4* __restrict__ C,
                                                                    const float4* __restrict__ Qr, const float4* __restrict__ Qs,
                                                                    float4* __restrict__ agg) {
  const int slot = threadIdx.x / (FP / 4), j = threadIdx.x - slot * (FP / 4);
  const int64_t r = (int64_t)blockIdx.x * TA_NODES + slot;
  if (r >= rows) return;
  const int64_t beg = row_ptr[r], end = min((int64_t)row_ptr[r + 1], E_cap), gb = (r / N) * N;
  const float4 qr = Qr[r * (FP / 4) + j];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t e = beg; e < end; ++e) {
    const float4 p = pre3(C[e * (FP / 4) + j], qr, Qs[(gb + send[e]) * (FP / 4) + j]);
    acc.x += fmaxf(p.x, 0.f); acc.y += fmaxf(p.y, 0.f); acc.z += fmaxf(p.z, 0.f); acc.w += fmaxf(p.w, 0.f);
  }
  agg[r * (FP / 4) + j] = acc;
}

// float4 j (columns 4j .. 4j+3) of row `row` of a feature buffer in either layout (blocked: columns 152..159 are not stored, zero)
__device__ __forceinline__ float4 feat4(const float* base, int64_t row, int j, bool blocked) {
  if (!blocked) return *reinterpret_cast<const float4*>(base + row * FP + 4 * j);
  if (4 * j >= tc::BLK_COLS) return make_float4(0.f, 0.f, 0.f, 0.f);
  return *reinterpret_cast<const float4*>(base + tc::blk_off(row, 4 * j));
}

// receiver side of the backward: with d = d_agg[n] (*) [pre_e > 0]:  dC[e] += d ;  dQr[n] = sum_{e in row n} d
__global__ void __launch_bounds__(TA_THREADS) effect_bwd_recv_kernel(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ send,
                                                                     int64_t rows, int N, int64_t E_cap, const float* __restrict__ C,
                                                                     const float* __restrict__ Qr, const float* __restrict__ Qs, bool blocked,
                                                                     const float4* __restrict__ d_agg, float4* __restrict__ dC, bool first,
                                                                     float4* __restrict__ dQr, float* __restrict__ qr_max) {
  __shared__ int smx[TA_NODES];   // per-row max |dQr| (bit pattern of a non-negative float orders like the float)
  if (threadIdx.x < TA_NODES) smx[threadIdx.x] = 0;
  __syncthreads();
  const int slot = threadIdx.x / (FP / 4), j = threadIdx.x - slot * (FP / 4);
  const int64_t r = (int64_t)blockIdx.x * TA_NODES + slot;
  if (r < rows) {
    const int64_t beg = row_ptr[r], end = min((int64_t)row_ptr[r + 1], E_cap), gb = (r / N) * N;
    const float4 qr = feat4(Qr, r, j, blocked), da = d_agg[r * (FP / 4) + j];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t e = beg; e < end; ++e) {
      const float4 p = pre3(feat4(C, e, j, blocked), qr, feat4(Qs, gb + send[e], j, blocked));
      const float4 d = make_float4(p.x > 0.f ? da.x : 0.f, p.y > 0.f ? da.y : 0.f, p.z > 0.f ? da.z : 0.f, p.w > 0.f ? da.w : 0.f);
      float4 c = d;
      if (!first) {   // the first propagation step visited (the last of the forward) initialises dC
        c = dC[e * (FP / 4) + j];
        c.x += d.x; c.y += d.y; c.z += d.z; c.w += d.w;
      }
      dC[e * (FP / 4) + j] = c;
      acc.x += d.x; acc.y += d.y; acc.z += d.z; acc.w += d.w;
    }
    dQr[r * (FP / 4) + j] = acc;
    if (qr_max) atomicMax(&smx[slot], __float_as_int(fmaxf(fmaxf(fabsf(acc.x), fabsf(acc.y)), fmaxf(fabsf(acc.z), fabsf(acc.w)))));
  }
  __syncthreads();
  if (qr_max && r < rows && j == 0) qr_max[r] = __int_as_float(smx[slot]);
}

// sender side: dQs[s] = sum over the relations sent by s (sender-sorted list) of d_agg[recv e] (*) [pre_e > 0]
__global__ void __launch_bounds__(TA_THREADS) effect_bwd_send_kernel(const int32_t* __restrict__ send_ptr, const int32_t* __restrict__ send_perm,
                                                                     const int32_t* __restrict__ recv, int64_t rows, int64_t E_cap,
                                                                     const float* __restrict__ C, const float* __restrict__ Qr,
                                                                     const float* __restrict__ Qs, bool blocked, const float4* __restrict__ d_agg,
                                                                     float4* __restrict__ dQs, float* __restrict__ qs_max) {
  __shared__ int smx[TA_NODES];
  if (threadIdx.x < TA_NODES) smx[threadIdx.x] = 0;
  __syncthreads();
  const int slot = threadIdx.x / (FP / 4), j = threadIdx.x - slot * (FP / 4);
  const int64_t s = (int64_t)blockIdx.x * TA_NODES + slot;
  if (s < rows) {
    const int64_t beg = send_ptr[s], end = min((int64_t)send_ptr[s + 1], E_cap);
    const float4 qs = feat4(Qs, s, j, blocked);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t i = beg; i < end; ++i) {
      const int e = send_perm[i];
      const int r = recv[e];
      const float4 p = pre3(feat4(C, e, j, blocked), feat4(Qr, r, j, blocked), qs);
      const float4 da = d_agg[(int64_t)r * (FP / 4) + j];
      acc.x += p.x > 0.f ? da.x : 0.f; acc.y += p.y > 0.f ? da.y : 0.f; acc.z += p.z > 0.f ? da.z : 0.f; acc.w += p.w > 0.f ? da.w : 0.f;
    }
    dQs[s * (FP / 4) + j] = acc;
    if (qs_max) atomicMax(&smx[slot], __float_as_int(fmaxf(fmaxf(fabsf(acc.x), fabsf(acc.y)), fmaxf(fabsf(acc.z), fabsf(acc.w)))));
  }
  __syncthreads();
  if (qs_max && s < rows && j == 0) qs_max[s] = __int_as_float(smx[slot]);
}

// ------------------------------------------------------------------------------------ head and state gradients
// motion = u2 * V2^T + c2 ; pred_pos = pos + clamp(motion)   (rows = all nodes; only n < n_p are written)
__global__ void head_out_kernel(const float* __restrict__ u2, const float* __restrict__ w2 /*[3][160] + 4 bias*/, const float* __restrict__ state,
                                int B, int N, int n_p, float* __restrict__ pred_pos, int64_t pos_stride_b, float* __restrict__ pred_motion) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= (int64_t)B * N) return;
  const int b = (int)(r / N), n = (int)(r - (int64_t)b * N);
  if (n >= n_p) return;
  const float* x = u2 + r * FP;
  float m[3];
#pragma unroll
  for (int o = 0; o < 3; ++o) {
    float s = 0.f;
    for (int k = 0; k < FP; ++k) s = fmaf(x[k], w2[o * FP + k], s);
    m[o] = s + w2[3 * FP + o];
  }
  const float* cur = state + (((size_t)b * H_FIX + (H_FIX - 1)) * N + n) * 3;
  float* mo = pred_motion + ((size_t)b * n_p + n) * 3;
  float* po = pred_pos + (size_t)b * pos_stride_b + (size_t)n * 3;
#pragma unroll
  for (int o = 0; o < 3; ++o) {
    mo[o] = m[o];
    po[o] = cur[o] + fminf(fmaxf(m[o], -MOTION_CLAMP_T), MOTION_CLAMP_T);
  }
}

// d_m = d_motion + d_pos * [|motion| <= clamp] ; d_u2[r][k] = sum_o d_m[o] * V2[o][k] (masked later by u2 > 0) ;
// dV2 / dc2 partial sums are left to a small dedicated reduction (head_wgrad_kernel)
__global__ void head_bwd_kernel(const float* __restrict__ d_pos, const float* __restrict__ d_motion, const float* __restrict__ motion,
                                const float* __restrict__ w2, int B, int N, int n_p, float* __restrict__ d_m_out /*[R][4]*/,
                                float* __restrict__ d_u2 /*[R][160]*/) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= (int64_t)B * N) return;
  const int b = (int)(r / N), n = (int)(r - (int64_t)b * N);
  float dm[3] = {0.f, 0.f, 0.f};
  if (n < n_p) {
    const size_t o3 = ((size_t)b * n_p + n) * 3;
#pragma unroll
    for (int o = 0; o < 3; ++o) {
      const float mo = motion[o3 + o];
      dm[o] = (d_motion ? d_motion[o3 + o] : 0.f) + ((d_pos && fabsf(mo) <= MOTION_CLAMP_T) ? d_pos[o3 + o] : 0.f);
    }
  }
  d_m_out[r * 4 + 0] = dm[0]; d_m_out[r * 4 + 1] = dm[1]; d_m_out[r * 4 + 2] = dm[2]; d_m_out[r * 4 + 3] = 0.f;
  float* du = d_u2 + r * FP;
  for (int k = 0; k < FP; ++k) du[k] = dm[0] * w2[k] + dm[1] * w2[FP + k] + dm[2] * w2[2 * FP + k];
}

// dV2[o][k] += sum_r d_m[r][o] * u2[r][k] ; dc2[o] += sum_r d_m[r][o]     (one block per (o, k-chunk), fixed order)
__global__ void __launch_bounds__(256) head_wgrad_kernel(const float* __restrict__ d_m, const float* __restrict__ u2, int64_t rows, int F,
                                                          float* __restrict__ dV2, float* __restrict__ dc2) {
  __shared__ float red[256];
  const int o = blockIdx.y, k = blockIdx.x;     // k in [0, F] ; k == F computes the bias gradient
  float s = 0.f;
  for (int64_t r = threadIdx.x; r < rows; r += 256) s += d_m[r * 4 + o] * (k < F ? u2[r * FP + k] : 1.f);
  red[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (k < F) dV2[o * F + k] += red[0];
    else dc2[o] += red[0];
  }
}

// d_rel_in (E x 24) -> d_nfeat: receiver side (+) over the CSR row, sender side (-) over the sender-sorted list; then
// d_state (B,H,N,3) += history-feature chain rule (+ d_pos on the last frame of object particles)
__global__ void state_bwd_kernel(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ send_ptr, const int32_t* __restrict__ send_perm,
                                 int64_t E_cap, const float* __restrict__ d_rel_in, const float* __restrict__ d_pos, int B, int N, int n_p,
                                 float* __restrict__ d_state) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= (int64_t)B * N) return;
  const int b = (int)(r / N), n = (int)(r - (int64_t)b * N);
  float dh[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) dh[i] = 0.f;
  for (int64_t e = row_ptr[r], end = min((int64_t)row_ptr[r + 1], E_cap); e < end; ++e) {
#pragma unroll
    for (int i = 0; i < 12; ++i) dh[i] += d_rel_in[e * D_REL_IN + 5 + i];
  }
  for (int64_t i0 = send_ptr[r], end = min((int64_t)send_ptr[r + 1], E_cap); i0 < end; ++i0) {
    const int64_t e = send_perm[i0];
#pragma unroll
    for (int i = 0; i < 12; ++i) dh[i] -= d_rel_in[e * D_REL_IN + 5 + i];
  }
  // hist = [s1-s0, s2-s1, s3-s2, s3]
  float ds[H_FIX][3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    ds[0][c] = -dh[c];
    ds[1][c] = dh[c] - dh[3 + c];
    ds[2][c] = dh[3 + c] - dh[6 + c];
    ds[3][c] = dh[6 + c] + dh[9 + c];
    if (d_pos && n < n_p) ds[3][c] += d_pos[((size_t)b * n_p + n) * 3 + c];   // pred_pos = state[:, -1, :n_p] + ...
  }
#pragma unroll
  for (int h = 0; h < H_FIX; ++h) {
    float* p = d_state + (((size_t)b * H_FIX + h) * N + n) * 3;
    p[0] += ds[h][0]; p[1] += ds[h][1]; p[2] += ds[h][2];
  }
}

__global__ void add_rows_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t n) {   // dst += src
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += src[i];
}
__global__ void mask_rows_kernel(float* __restrict__ dst, const float* __restrict__ act, int64_t n) {  // dst *= (act > 0)
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = act[i] > 0.f ? dst[i] : 0.f;
}

// ------------------------------------------------------------------------------------ packed weights for the training path
// Forward uses the k-major copies of PackedLayout; dgrad needs the plain [n][j] (padded to [160][160]) copies:
struct TrainLayout {
  size_t penc2, penc4, renc0, renc2, renc4, rp_rel, rp_recv, rp_send, pp_enc, pp_agg, pred0, pred1, total;   // float offsets
};
TrainLayout train_layout(size_t base_floats) {
  TrainLayout L;
  size_t o = align_up(base_floats, 64);
  auto m = [&]() { size_t r = o; o += (size_t)FP * FP; return r; };
  L.penc2 = m(); L.penc4 = m(); L.renc0 = m(); L.renc2 = m(); L.renc4 = m(); L.rp_rel = m(); L.rp_recv = m(); L.rp_send = m();
  L.pp_enc = m(); L.pp_agg = m(); L.pred0 = m(); L.pred1 = m();
  L.total = o;
  return L;
}
// dst[n*160 + j] = W[n*ld + col0 + j]  (n < F, j < K), zero padded: the k-major operand of the dgrad product dX = dY * W
__global__ void pack_plain_kernel(const float* __restrict__ W, int ld, int col0, int K, int F, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= FP * FP) return;
  const int n = i / FP, j = i - n * FP;
  dst[i] = (n < F && j < K) ? W[(size_t)n * ld + col0 + j] : 0.f;
}

}  // namespace agx

// ===================================================================================== host orchestration
namespace agx {

size_t tc_blob_bytes(size_t base_bytes);

struct TrainSaved {
  float *nfeat, *p_in, *rel_in, *h1, *h2, *penc, *g1, *g2, *renc, *A, *C, *u1, *u2, *motion;
  float *P[17], *agg[16], *Qr[16], *Qs[16];   // P[0] aliases penc
  // tensor-core forward (default): C, Qr[k], Qs[k] hold the blocked layout of tc_chain.cuh (only the relation kernels read them);
  // the chains' own streams follow
  float *A_blk, *P_blk, *P0_blk, *S0, *rowmaxP, *rowmaxA, *rowmaxP0, *agg_split, *agg_max;
  int32_t* agg_exp;
};
// floats of a [n][FP] feature buffer that may hold either layout
static size_t feat_floats(int64_t n) {
  const size_t rm = (size_t)n * FP, bl = (size_t)tc::blk_rows(n) * tc::BLK_COLS;
  return rm > bl ? rm : bl;
}
static size_t saved_carve(void* base, int64_t rows, int64_t E, int K, TrainSaved* out) {
  Carver c(base);
  TrainSaved s;
  s.nfeat = c.take<float>(rows * NFEAT); s.p_in = c.take<float>(rows * D_NODE_IN); s.rel_in = c.take<float>(E * D_REL_IN);
  s.h1 = c.take<float>(rows * FP); s.h2 = c.take<float>(rows * FP); s.penc = c.take<float>(rows * FP);
  s.g1 = c.take<float>(E * FP); s.g2 = c.take<float>(E * FP); s.renc = c.take<float>(E * FP);
  s.A = c.take<float>(rows * FP); s.C = c.take<float>(feat_floats(E));
  s.u1 = c.take<float>(rows * FP); s.u2 = c.take<float>(rows * FP); s.motion = c.take<float>(rows * 3);
  s.P[0] = s.penc;
  for (int k = 0; k < K; ++k) {
    s.P[k + 1] = c.take<float>(rows * FP); s.agg[k] = c.take<float>(rows * FP);
    s.Qr[k] = c.take<float>(feat_floats(rows)); s.Qs[k] = c.take<float>(feat_floats(rows));
  }
  const size_t blk = (size_t)tc::blk_rows(rows) * tc::BLK_COLS;
  s.A_blk = c.take<float>(blk); s.P_blk = c.take<float>(blk); s.P0_blk = c.take<float>(blk); s.S0 = c.take<float>(blk);
  s.rowmaxP = c.take<float>(rows); s.rowmaxA = c.take<float>(rows); s.rowmaxP0 = c.take<float>(rows);
  s.agg_split = c.take<float>(blk); s.agg_max = c.take<float>(rows); s.agg_exp = c.take<int32_t>(rows);
  if (out) *out = s;
  return align_up(c.off, 256);
}

struct TrainScratch {
  float *dU, *dV, *dH2, *dH1, *dP, *dPn, *dA, *dAgg, *dQr, *dQs, *dC, *dE, *dG2, *dG1, *dRel, *dm, *part;
  float *preMax, *aBound, *aggMax, *cBound, *qrMax, *qsMax;   // per-row magnitude bounds of the tensor-core backward
  float *dPreK[16], *dQrK[16], *dQsK[16], *dP0;   // per propagation step: kept until the single weight-gradient batch at the end
};
static int wgrad_max_ctas() { return num_sms() + tc::WG_MAX_JOBS; }
static size_t scratch_carve(void* base, int64_t rows, int64_t E, int K, TrainScratch* out) {
  Carver c(base);
  TrainScratch s;
  s.dU = c.take<float>(rows * FP); s.dV = c.take<float>(rows * FP); s.dH2 = c.take<float>(rows * FP); s.dH1 = c.take<float>(rows * FP);
  s.dP = c.take<float>(rows * FP); s.dPn = c.take<float>(rows * FP);
  s.dA = c.take<float>(rows * FP); s.dAgg = c.take<float>(rows * FP); s.dQr = c.take<float>(rows * FP); s.dQs = c.take<float>(rows * FP);
  s.dC = c.take<float>(E * FP); s.dE = c.take<float>(E * FP); s.dG2 = c.take<float>(E * FP); s.dG1 = c.take<float>(E * FP);
  s.dRel = c.take<float>(E * D_REL_IN); s.dm = c.take<float>(rows * 4);
  const size_t part_simt = (size_t)num_sms() * (FP * FP + FP), part_tc = tc_wgrad_part_floats(wgrad_max_ctas());
  s.part = c.take<float>(part_simt > part_tc ? part_simt : part_tc);
  s.preMax = c.take<float>(rows); s.aBound = c.take<float>(rows); s.aggMax = c.take<float>(rows); s.cBound = c.take<float>(rows);
  s.qrMax = c.take<float>(rows); s.qsMax = c.take<float>(rows);
  for (int k = 0; k < K; ++k) { s.dPreK[k] = c.take<float>(rows * FP); s.dQrK[k] = c.take<float>(rows * FP); s.dQsK[k] = c.take<float>(rows * FP); }
  s.dP0 = c.take<float>(rows * FP);
  if (out) *out = s;
  return align_up(c.off, 256);
}

static size_t train_base_floats() { return tc_blob_bytes(packed_layout().total * sizeof(float)) / sizeof(float); }

static int train_attrs() {
  static thread_local DeviceOnce once;
  if (!once.need()) return AGX_OK;
  AGX_CUDA_OK(cudaFuncSetAttribute(lin_kernel<D_NODE_IN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MLP_SMEM_BYTES));
  AGX_CUDA_OK(cudaFuncSetAttribute(lin_kernel<D_REL_IN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MLP_SMEM_BYTES));
  AGX_CUDA_OK(cudaFuncSetAttribute(lin_kernel<FP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MLP_SMEM_BYTES));
  AGX_CUDA_OK(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WG_SMEM));
  return AGX_OK;
}

template <int K>
static int lin(cudaStream_t st, const float* X, int ldx, const float* mask, int ldm, const float* Wt, const float* bias, const float* add1,
               const float* add2, float* Y, int ldy, int64_t M, bool relu, bool accumulate, int n_store = FP) {
  if (M <= 0) return AGX_OK;
  LinArgs a{X, ldx, mask, ldm, Wt, bias, add1, add2, Y, ldy, M, relu ? 1 : 0, accumulate ? 1 : 0, n_store};
  const int64_t tiles = (M + TM - 1) / TM;
  const int grid = (int)(tiles < 2 * num_sms() ? tiles : 2 * num_sms());
  { ProfScope ps(AGX_KIND_OTHER, st);
    lin_kernel<K><<<grid, MLP_THREADS, MLP_SMEM_BYTES, st>>>(a); }
  AGX_LAUNCH_CHECK();
  return AGX_OK;
}

// 160 -> 160 layers: exact fp32 FFMA tiles (above) by default; AGX_TRAIN_PRECISION=tc runs them on the tensor cores
// (tc_forward.cu: tc_lin_kernel).  The split-fp16 products carry 22 significant bits per operand: ample for the forward pass
// (positive post-ReLU activations), but back-propagated gradients cancel heavily and the error compounds along the chain —
// measured on granular-150 x 3, pstep 2 against torch autograd: worst parameter gradient 1.4e-4 of its maximum with the tensor-core
// layers, 8e-7 with the FFMA tiles (tests/bench/grad_check.py) — beyond the 1e-4 the parity tests state, hence opt-in.
// `layer` names the fp16 image of the same matrix inside `packed`.
int tc_lin(cudaStream_t st, const void* packed, size_t base_bytes, int layer, const float* X, int ldx, const float* mask, int ldm,
           const float* bias, const float* add1, const float* add2, float* Y, int ldy, int64_t M, bool relu, bool accumulate, int n_store);
static int train_use_tc() {   // 0 fp32, 1 every 160x160 layer, 2 forward layers only, 3 backward (dgrad) layers only
  static const int v = [] {
    const char* e = getenv("AGX_TRAIN_PRECISION");
    if (!e) return 0;
    return !strcmp(e, "tc") ? 1 : !strcmp(e, "tc_fwd") ? 2 : !strcmp(e, "tc_bwd") ? 3 : !strcmp(e, "tc_one") ? 4 : 0;
  }();
  return v;
}
// forward of the training step: the tensor-core chains of the inference path with their training copies (default), or the
// layer-by-layer fp32 FFMA kernels of this file (AGX_TRAIN_FORWARD=fp32)
static bool train_forward_tc() {
  static const bool v = [] { const char* e = getenv("AGX_TRAIN_FORWARD"); return !(e && !strcmp(e, "fp32")); }();
  return v;
}
// backward: chain kernels on the tensor cores (tc_backward.inl, default) or the layer-by-layer path below (AGX_TRAIN_BACKWARD=fp32)
static bool train_backward_tc() {
  static const bool v = [] { const char* e = getenv("AGX_TRAIN_BACKWARD"); return !(e && !strcmp(e, "fp32")); }();
  return v;
}
static int train_tc_layer() {
  static const int v = [] { const char* e = getenv("AGX_TRAIN_TC_LAYER"); return e ? atoi(e) : -1; }();
  return v;
}
static int lin160(cudaStream_t st, const float* packed, int layer, const float* X, int ldx, const float* mask, int ldm, const float* Wt,
                  const float* bias, const float* add1, const float* add2, float* Y, int ldy, int64_t M, bool relu, bool accumulate,
                  int n_store = FP) {
  const int mode = train_use_tc();
  if (mode == 1 || (mode == 2 && layer < tc::TT_PENC2) || (mode == 3 && layer >= tc::TT_PENC2) || (mode == 4 && layer == train_tc_layer()))
    return tc_lin(st, packed, packed_layout().total * sizeof(float), layer, X, ldx, mask, ldm, bias, add1, add2, Y, ldy, M, relu, accumulate, n_store);
  return lin<FP>(st, X, ldx, mask, ldm, Wt, bias, add1, add2, Y, ldy, M, relu, accumulate, n_store);
}

// Weight gradients are collected into batches: on the tensor cores (default) a batch is one launch of tc_wgrad.cu's kernel plus
// its reduction; with AGX_TRAIN_PRECISION=fp32 every job runs at once on the FFMA kernel above.
static bool train_wgrad_tc() {
  static const bool v = [] { const char* e = getenv("AGX_TRAIN_WGRAD"); return !(e && !strcmp(e, "fp32")); }();
  return v;
}
static int wgrad(cudaStream_t st, const float* dY, const float* mask, const float* X, int ldx, int kx, int64_t M, float* part, int F, int K,
                 int ld, int col0, float* dW, float* db);
struct WgradBatch {
  tc::WgArgs a;
  cudaStream_t st;
  float* part;
  explicit WgradBatch(cudaStream_t s, float* p) : st(s), part(p) { a.njobs = 0; a.part = nullptr; }
  int add(const float* dY, const float* mask, const float* X, int ldx, int kx, int64_t M, int F, int K, int ld, int col0, float* dW, float* db,
          const int32_t* m_limit = nullptr) {
    if (M <= 0 || !dW) return AGX_OK;
    if (!train_wgrad_tc()) return wgrad(st, dY, mask, X, ldx, kx, M, part, F, K, ld, col0, dW, db);
    if (a.njobs == tc::WG_MAX_JOBS) { if (int rc = flush()) return rc; }
    tc::WgJob& j = a.job[a.njobs++];
    j.dY = dY; j.mask = mask; j.X = X; j.dW = dW; j.db = db; j.M = M; j.m_limit = m_limit; j.ldx = ldx; j.kx = kx;
    j.npad = kx >= FP ? FP : (kx + 1 + 15) / 16 * 16;
    j.bias_col = db ? 1 : 0;
    j.ld = ld; j.col0 = col0; j.F = F; j.K = K; j.chain_head = 1; j.next = -1;
    for (int i = a.njobs - 2; i >= 0; --i)   // same destination as an earlier job of the batch: summed by that job's reduction
      if (a.job[i].dW == dW && a.job[i].col0 == col0 && a.job[i].next < 0) { a.job[i].next = a.njobs - 1; j.chain_head = 0; j.db = nullptr; break; }
    return AGX_OK;
  }
  int flush() {
    if (a.njobs == 0) return AGX_OK;
    const int rc = tc_wgrad_batch(st, a, part, wgrad_max_ctas());
    a.njobs = 0;
    return rc;
  }
};

static int wb_hold(WgradBatch& wb, const float* dY, const float* mask, const float* X, int ldx, int kx, int64_t M, int F, int K, int ld,
                   int col0, float* dW, float* db, const int32_t* m_limit = nullptr) {
  return wb.add(dY, mask, X, ldx, kx, M, F, K, ld, col0, dW, db, m_limit);
}

// dW (reference layout, += ) and optional db from dY (masked by act > 0) and the layer input X
static int wgrad(cudaStream_t st, const float* dY, const float* mask, const float* X, int ldx, int kx, int64_t M, float* part, int F, int K,
                 int ld, int col0, float* dW, float* db) {
  if (M <= 0 || !dW) return AGX_OK;
  const int64_t tiles = (M + TM - 1) / TM;
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  WgradArgs a{dY, FP, mask, FP, X, ldx, kx, M, part};
  { ProfScope ps(AGX_KIND_OTHER, st);
    wgrad_kernel<<<grid, MLP_THREADS, WG_SMEM, st>>>(a); }
  AGX_LAUNCH_CHECK();
  const int n = F * K + (db ? F : 0);
  { ProfScope ps(AGX_KIND_OTHER, st);
    wgrad_reduce_kernel<<<(n + 255) / 256, 256, 0, st>>>(part, grid, F, K, ld, col0, dW, db); }
  AGX_LAUNCH_CHECK();
  return AGX_OK;
}

#define AGX_TRY(expr) do { if (int _rc = (expr)) return _rc; } while (0)

int train_pack(const AgxModelDims* dims, const AgxWeights* raw, void* packed, cudaStream_t st) {
  const TrainLayout T = train_layout(train_base_floats());
  float* out = static_cast<float*>(packed);
  const int F = dims->F;
  const int d_rel = 2 * dims->d_attr + 1 + 3 * dims->n_his;
  auto plain = [&](int layer, int ld, int col0, int K, size_t off) -> int {
    pack_plain_kernel<<<(FP * FP + 255) / 256, 256, 0, st>>>(raw->weight[layer], ld, col0, K, F, out + off);
    AGX_LAUNCH_CHECK();
    return AGX_OK;
  };
  AGX_TRY(plain(AGX_W_PENC2, F, 0, F, T.penc2)); AGX_TRY(plain(AGX_W_PENC4, F, 0, F, T.penc4));
  AGX_TRY(plain(AGX_W_RENC0, d_rel, 0, d_rel, T.renc0)); AGX_TRY(plain(AGX_W_RENC2, F, 0, F, T.renc2)); AGX_TRY(plain(AGX_W_RENC4, F, 0, F, T.renc4));
  AGX_TRY(plain(AGX_W_RPROP, 3 * F, 0, F, T.rp_rel)); AGX_TRY(plain(AGX_W_RPROP, 3 * F, F, F, T.rp_recv)); AGX_TRY(plain(AGX_W_RPROP, 3 * F, 2 * F, F, T.rp_send));
  AGX_TRY(plain(AGX_W_PPROP, 2 * F, 0, F, T.pp_enc)); AGX_TRY(plain(AGX_W_PPROP, 2 * F, F, F, T.pp_agg));
  AGX_TRY(plain(AGX_W_PRED0, F, 0, F, T.pred0)); AGX_TRY(plain(AGX_W_PRED1, F, 0, F, T.pred1));
  return AGX_OK;
}
size_t train_blob_bytes() { return train_layout(train_base_floats()).total * sizeof(float); }

}  // namespace agx

extern "C" {

size_t agx_train_saved_bytes(const AgxModelDims* dims, int32_t B, int32_t N, int64_t E_cap) {
  if (!dims || B <= 0 || N <= 0 || E_cap < 0 || dims->pstep < 1 || dims->pstep > 16) return 0;
  return agx::saved_carve(nullptr, (int64_t)B * N, E_cap > 0 ? E_cap : 1, dims->pstep, nullptr);
}
size_t agx_train_scratch_bytes(const AgxModelDims* dims, int32_t B, int32_t N, int64_t E_cap) {
  if (!dims || B <= 0 || N <= 0 || E_cap < 0) return 0;
  if (dims->pstep < 1 || dims->pstep > 16) return 0;
  return agx::scratch_carve(nullptr, (int64_t)B * N, E_cap > 0 ? E_cap : 1, dims->pstep, nullptr);
}

int agx_train_saved_offsets(const AgxModelDims* dims, int32_t B, int32_t N, int64_t E_cap, int64_t* out, int32_t n) {
  using namespace agx;
  AGX_REQUIRE(dims && out && B > 0 && N > 0 && E_cap >= 0 && dims->pstep >= 1 && dims->pstep <= 16, AGX_ERR_ARG, "train_saved_offsets: bad argument");
  AGX_REQUIRE(n >= 9 + 4 * dims->pstep, AGX_ERR_CAPACITY, "train_saved_offsets: %d entries < %d", n, 9 + 4 * dims->pstep);
  TrainSaved s;
  saved_carve(nullptr, (int64_t)B * N, E_cap > 0 ? E_cap : 1, dims->pstep, &s);
  auto off = [](const float* p) { return (int64_t)reinterpret_cast<intptr_t>(p); };   // carved from a null base: the pointer is the offset
  const float* fixed[9] = {s.h1, s.h2, s.penc, s.g1, s.g2, s.renc, s.C, s.u1, s.u2};
  for (int i = 0; i < 9; ++i) out[i] = off(fixed[i]);
  for (int k = 0; k < dims->pstep; ++k) {
    out[9 + 4 * k] = off(s.P[k + 1]); out[10 + 4 * k] = off(s.agg[k]); out[11 + 4 * k] = off(s.Qr[k]); out[12 + 4 * k] = off(s.Qs[k]);
  }
  return train_forward_tc() ? 1 : 0;
}

int agx_forward_train(const AgxModelDims* dims, const void* packed_weights, const AgxGraphIn* g, float* pred_pos, int64_t pos_stride_b,
                      float* pred_motion, void* saved, size_t saved_bytes, agx_stream_t stream) {
  using namespace agx;
  AGX_REQUIRE(dims && packed_weights && g && pred_pos && pred_motion && saved, AGX_ERR_ARG, "forward_train: null pointer argument");
  AGX_REQUIRE(dims->pstep >= 1 && dims->pstep <= 16 && dims->n_his == H_FIX && dims->d_attr == 2 && dims->d_phys == 1 && dims->d_act == 3,
              AGX_ERR_ARG, "forward_train: unsupported model dims");
  AGX_REQUIRE(g->B > 0 && g->N > 0 && g->n_p > 0 && g->n_p <= g->N && g->E_cap >= 0, AGX_ERR_ARG, "forward_train: bad graph sizes");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t rows = (int64_t)g->B * g->N, E = g->E_cap > 0 ? g->E_cap : 1;
  const int K = dims->pstep;
  TrainSaved s;
  const size_t need = saved_carve(saved, rows, E, K, &s);
  AGX_REQUIRE(saved_bytes >= need, AGX_ERR_CAPACITY, "forward_train: saved buffer %zu < %zu bytes", saved_bytes, need);
  AGX_TRY(train_attrs());
  const PackedLayout L = packed_layout();
  const float* W = static_cast<const float*>(packed_weights);
  if (train_forward_tc()) {
    // the inference chains (tc_forward.cu, three fp16 products per term, fp32 C), each epilogue also leaving the fp32 row the
    // backward reads; per propagation step its own Qr / Qs (blocked) and P / agg (row-major) buffers
    const size_t base = L.total * sizeof(float);
    TcTrainSave sv{s.p_in, s.h1, s.h2, s.penc, s.rel_in, s.g1, s.g2, s.renc, nullptr, nullptr, s.u1, s.u2};
    TcFwdBuffers tb{s.nfeat, s.P_blk, s.A_blk, nullptr, nullptr, s.agg_split, s.C, s.rowmaxP, s.rowmaxA, s.agg_exp, s.agg_max,
                    s.P0_blk, s.Qr[0], s.Qs[0], s.rowmaxP0, s.S0, &sv};
    AGX_TRY(tc_node_encoder(g, W, L, base, tb, st));
    if (g->E_cap > 0) AGX_TRY(tc_edge_encoder(g, W, L, base, tb, false, st));
    for (int k = 0; k < K; ++k) {
      sv.agg_f32 = s.agg[k];
      sv.P_next = s.P[k + 1];
      tb.Qr = s.Qr[k]; tb.Qs = s.Qs[k];                       // read by the aggregate of step k > 0
      AGX_TRY(tc_edge_aggregate(g, tb, false, k == 0, st));
      if (k + 1 < K) { tb.Qr = s.Qr[k + 1]; tb.Qs = s.Qs[k + 1]; }   // written by the update of step k
      AGX_TRY(tc_node_update(g, W, L, base, tb, k == 0, k + 1 == K, pred_pos, pos_stride_b, pred_motion, st));
    }
    return AGX_OK;
  }
  node_prep_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(g->state, g->attrs, g->action, g->p_instance, g->physics, g->B, g->N, g->n_p,
                                                                   s.nfeat, s.p_in);
  AGX_LAUNCH_CHECK();
  AGX_TRY(lin<D_NODE_IN>(st, s.p_in, D_NODE_IN, nullptr, 0, W + L.penc0_w, W + L.penc0_b, nullptr, nullptr, s.h1, FP, rows, true, false));
  AGX_TRY(lin160(st, W, tc::T_PENC2, s.h1, FP, nullptr, 0, W + L.penc2_w, W + L.penc2_b, nullptr, nullptr, s.h2, FP, rows, true, false));
  AGX_TRY(lin160(st, W, tc::T_PENC4, s.h2, FP, nullptr, 0, W + L.penc4_w, W + L.penc4_b, nullptr, nullptr, s.penc, FP, rows, true, false));
  AGX_TRY(lin160(st, W, tc::T_PP_ENC, s.penc, FP, nullptr, 0, W + L.pp_enc_w, W + L.pp_b, nullptr, nullptr, s.A, FP, rows, false, false));
  if (g->E_cap > 0) {
    edge_prep_kernel<<<(unsigned)((E + 255) / 256), 256, 0, st>>>(g->row_ptr, g->send, g->recv, rows, g->N, g->E_cap, s.nfeat, s.rel_in);
    AGX_LAUNCH_CHECK();
    AGX_TRY(lin<D_REL_IN>(st, s.rel_in, D_REL_IN, nullptr, 0, W + L.renc0_w, W + L.renc0_b, nullptr, nullptr, s.g1, FP, E, true, false));
    AGX_TRY(lin160(st, W, tc::T_RENC2, s.g1, FP, nullptr, 0, W + L.renc2_w, W + L.renc2_b, nullptr, nullptr, s.g2, FP, E, true, false));
    AGX_TRY(lin160(st, W, tc::T_RENC4, s.g2, FP, nullptr, 0, W + L.renc4_w, W + L.renc4_b, nullptr, nullptr, s.renc, FP, E, true, false));
    AGX_TRY(lin160(st, W, tc::T_RP_REL, s.renc, FP, nullptr, 0, W + L.rp_rel_w, W + L.rp_b, nullptr, nullptr, s.C, FP, E, false, false));
  }
  for (int k = 0; k < K; ++k) {
    AGX_TRY(lin160(st, W, tc::T_RP_RECV, s.P[k], FP, nullptr, 0, W + L.rp_recv_w, nullptr, nullptr, nullptr, s.Qr[k], FP, rows, false, false));
    AGX_TRY(lin160(st, W, tc::T_RP_SEND, s.P[k], FP, nullptr, 0, W + L.rp_send_w, nullptr, nullptr, nullptr, s.Qs[k], FP, rows, false, false));
    train_aggregate_kernel<<<(unsigned)((rows + TA_NODES - 1) / TA_NODES), TA_THREADS, 0, st>>>(
        g->row_ptr, g->send, rows, g->N, g->E_cap, reinterpret_cast<const float4*>(s.C), reinterpret_cast<const float4*>(s.Qr[k]),
        reinterpret_cast<const float4*>(s.Qs[k]), reinterpret_cast<float4*>(s.agg[k]));
    AGX_LAUNCH_CHECK();
    AGX_TRY(lin160(st, W, tc::T_PP_AGG, s.agg[k], FP, nullptr, 0, W + L.pp_agg_w, nullptr, s.A, s.P[k], s.P[k + 1], FP, rows, true, false));
  }
  AGX_TRY(lin160(st, W, tc::T_PRED0, s.P[K], FP, nullptr, 0, W + L.pred0_w, W + L.pred0_b, nullptr, nullptr, s.u1, FP, rows, true, false));
  AGX_TRY(lin160(st, W, tc::T_PRED1, s.u1, FP, nullptr, 0, W + L.pred1_w, W + L.pred1_b, nullptr, nullptr, s.u2, FP, rows, true, false));
  head_out_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, st>>>(s.u2, W + L.pred2_w, g->state, g->B, g->N, g->n_p, pred_pos, pos_stride_b, pred_motion);
  AGX_LAUNCH_CHECK();
  return AGX_OK;
}

int agx_backward(const AgxModelDims* dims, const void* packed_weights, const AgxGraphIn* g, const void* saved, const int32_t* send_ptr,
                 const int32_t* send_perm, const float* pred_motion, const float* d_pred_pos, const float* d_pred_motion,
                 const AgxWeightGrads* grads, float* d_state, void* scratch, size_t scratch_bytes, agx_stream_t stream) {
  using namespace agx;
  AGX_REQUIRE(dims && packed_weights && g && saved && grads && scratch && pred_motion, AGX_ERR_ARG, "backward: null pointer argument");
  AGX_REQUIRE(g->E_cap == 0 || (send_ptr && send_perm), AGX_ERR_ARG, "backward: sender-sorted relation lists are required");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t rows = (int64_t)g->B * g->N, E = g->E_cap > 0 ? g->E_cap : 1;
  const int K = dims->pstep, F = dims->F;
  const int d_node = dims->d_attr + dims->d_phys + dims->d_act, d_rel = 2 * dims->d_attr + 1 + 3 * dims->n_his;
  TrainSaved s;
  saved_carve(const_cast<void*>(saved), rows, E, K, &s);
  TrainScratch t;
  const size_t need = scratch_carve(scratch, rows, E, K, &t);
  AGX_REQUIRE(scratch_bytes >= need, AGX_ERR_CAPACITY, "backward: scratch %zu < %zu bytes", scratch_bytes, need);
  AGX_TRY(train_attrs());
  const PackedLayout L = packed_layout();
  const TrainLayout T = train_layout(train_base_floats());
  const float* W = static_cast<const float*>(packed_weights);
  float* const* gw = grads->weight;
  float* const* gb = grads->bias;
  const bool fwd_tc = train_forward_tc();   // layout of the saved C / Qr / Qs
  if (train_backward_tc()) {
    // Four kinds of chain kernels (tc_backward.cuh) + the relation kernels + batched weight-gradient jobs.  Every gradient row
    // leaves its chain masked, so the weight-gradient jobs read two streams (gradient, layer input) and no mask.
    const size_t base = L.total * sizeof(float);
    const TcBwdBuffers bb{s.u2, s.u1, s.penc, s.h2, s.h1, s.renc, s.g2, s.g1,
                          t.dm, t.dU, t.dV, t.dPreK[K - 1], nullptr, t.dA, t.dAgg, t.dQr, t.dQs, t.dP, t.dH2, t.dH1, t.dC, t.dE, t.dG2, t.dG1, t.dRel,
                          t.preMax, t.aBound, t.aggMax, t.cBound, t.qrMax, t.qsMax};
    WgradBatch wb(st, t.part);
    AGX_TRY(tc_bwd_head(g, W, L, base, bb, s.P[K], d_pred_pos, d_pred_motion, pred_motion, st));
    if (gw[AGX_W_PRED2]) {
      head_wgrad_kernel<<<dim3(F + 1, 3), 256, 0, st>>>(t.dm, s.u2, rows, F, gw[AGX_W_PRED2], gb[AGX_W_PRED2]);
      AGX_LAUNCH_CHECK();
    }
    AGX_TRY(wb.add(t.dU, nullptr, s.u1, FP, FP, rows, F, F, F, 0, gw[AGX_W_PRED1], gb[AGX_W_PRED1]));
    AGX_TRY(wb.add(t.dV, nullptr, s.P[K], FP, FP, rows, F, F, F, 0, gw[AGX_W_PRED0], gb[AGX_W_PRED0]));
    if (g->E_cap == 0) {
      for (int k = 0; k < K; ++k) {
        AGX_CUDA_OK(cudaMemsetAsync(t.dQrK[k], 0, rows * FP * sizeof(float), st));
        AGX_CUDA_OK(cudaMemsetAsync(t.dQsK[k], 0, rows * FP * sizeof(float), st));
      }
      AGX_CUDA_OK(cudaMemsetAsync(t.qrMax, 0, rows * sizeof(float), st));
      AGX_CUDA_OK(cudaMemsetAsync(t.qsMax, 0, rows * sizeof(float), st));
    }
    for (int k = K - 1; k >= 0; --k) {
      // every step keeps its own d pre / dQr / dQs rows, so that all weight-gradient jobs run as ONE batch at the end
      TcBwdBuffers bk = bb;
      bk.dPre = t.dPreK[k]; bk.dQr = t.dQrK[k]; bk.dQs = t.dQsK[k];
      bk.dPreOut = k > 0 ? t.dPreK[k - 1] : t.dP0;
      AGX_TRY(wb.add(bk.dPre, nullptr, s.agg[k], FP, FP, rows, F, F, 2 * F, F, gw[AGX_W_PPROP], nullptr));
      if (g->E_cap > 0) {
        effect_bwd_recv_kernel<<<(unsigned)((rows + TA_NODES - 1) / TA_NODES), TA_THREADS, 0, st>>>(
            g->row_ptr, g->send, rows, g->N, g->E_cap, s.C, s.Qr[k], s.Qs[k], fwd_tc, reinterpret_cast<const float4*>(t.dAgg),
            reinterpret_cast<float4*>(t.dC), k == K - 1, reinterpret_cast<float4*>(bk.dQr), t.qrMax);
        AGX_LAUNCH_CHECK();
        effect_bwd_send_kernel<<<(unsigned)((rows + TA_NODES - 1) / TA_NODES), TA_THREADS, 0, st>>>(
            send_ptr, send_perm, g->recv, rows, g->E_cap, s.C, s.Qr[k], s.Qs[k], fwd_tc, reinterpret_cast<const float4*>(t.dAgg),
            reinterpret_cast<float4*>(bk.dQs), t.qsMax);
        AGX_LAUNCH_CHECK();
        AGX_TRY(wb.add(bk.dQr, nullptr, s.P[k], FP, FP, rows, F, F, 3 * F, F, gw[AGX_W_RPROP], nullptr));
        AGX_TRY(wb.add(bk.dQs, nullptr, s.P[k], FP, FP, rows, F, F, 3 * F, 2 * F, gw[AGX_W_RPROP], nullptr));
      }
      if (k > 0) AGX_TRY(tc_bwd_step(g, W, base, bk, s.P[k], st));
      else AGX_TRY(tc_bwd_node_encoder(g, W, base, bk, st));
    }
    AGX_TRY(wb.add(t.dA, nullptr, s.penc, FP, FP, rows, F, F, 2 * F, 0, gw[AGX_W_PPROP], gb[AGX_W_PPROP]));
    AGX_TRY(wb.add(t.dP, nullptr, s.h2, FP, FP, rows, F, F, F, 0, gw[AGX_W_PENC4], gb[AGX_W_PENC4]));
    AGX_TRY(wb.add(t.dH2, nullptr, s.h1, FP, FP, rows, F, F, F, 0, gw[AGX_W_PENC2], gb[AGX_W_PENC2]));
    AGX_TRY(wb.add(t.dH1, nullptr, s.p_in, D_NODE_IN, D_NODE_IN, rows, F, d_node, d_node, 0, gw[AGX_W_PENC0], gb[AGX_W_PENC0]));
    if (g->E_cap > 0) {
      AGX_TRY(tc_bwd_edge_encoder(g, W, base, bb, st));
      const int32_t* n_rel = g->row_ptr + rows;   // relations actually built (rows past it hold nothing)
      AGX_TRY(wb.add(t.dC, nullptr, s.renc, FP, FP, E, F, F, 3 * F, 0, gw[AGX_W_RPROP], gb[AGX_W_RPROP], n_rel));
      AGX_TRY(wb.add(t.dE, nullptr, s.g2, FP, FP, E, F, F, F, 0, gw[AGX_W_RENC4], gb[AGX_W_RENC4], n_rel));
      AGX_TRY(wb.add(t.dG2, nullptr, s.g1, FP, FP, E, F, F, F, 0, gw[AGX_W_RENC2], gb[AGX_W_RENC2], n_rel));
      AGX_TRY(wb.add(t.dG1, nullptr, s.rel_in, D_REL_IN, D_REL_IN, E, F, d_rel, d_rel, 0, gw[AGX_W_RENC0], gb[AGX_W_RENC0], n_rel));
    }
    AGX_TRY(wb.flush());
    if (d_state) {
      if (g->E_cap == 0) AGX_CUDA_OK(cudaMemsetAsync(t.dRel, 0, sizeof(float), st));
      state_bwd_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, st>>>(g->row_ptr, g->E_cap > 0 ? send_ptr : g->row_ptr, send_perm, g->E_cap, t.dRel,
                                                                      d_pred_pos, g->B, g->N, g->n_p, d_state);
      AGX_LAUNCH_CHECK();
    }
    return AGX_OK;
  }
  AGX_CUDA_OK(cudaMemsetAsync(t.dA, 0, rows * FP * sizeof(float), st));
  AGX_CUDA_OK(cudaMemsetAsync(t.dC, 0, E * FP * sizeof(float), st));

  // ---- head
  WgradBatch wb(st, t.part);
  head_bwd_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, st>>>(d_pred_pos, d_pred_motion, pred_motion, W + L.pred2_w, g->B, g->N, g->n_p, t.dm, t.dU);
  AGX_LAUNCH_CHECK();
  if (gw[AGX_W_PRED2]) {
    head_wgrad_kernel<<<dim3(F + 1, 3), 256, 0, st>>>(t.dm, s.u2, rows, F, gw[AGX_W_PRED2], gb[AGX_W_PRED2]);
    AGX_LAUNCH_CHECK();
  }
  // weight-gradient jobs whose operands stay untouched until the end of the backward (dU, dV here; the encoder gradients below)
  // wait for the final batch; the per-step ones are flushed before their operands are overwritten
  AGX_TRY(wb_hold(wb, t.dU, s.u2, s.u1, FP, FP, rows, F, F, F, 0, gw[AGX_W_PRED1], gb[AGX_W_PRED1]));
  AGX_TRY(lin160(st, W, tc::TT_PRED1, t.dU, FP, s.u2, FP, W + T.pred1, nullptr, nullptr, nullptr, t.dV, FP, rows, false, false));          // dU1
  AGX_TRY(wb_hold(wb, t.dV, s.u1, s.P[K], FP, FP, rows, F, F, F, 0, gw[AGX_W_PRED0], gb[AGX_W_PRED0]));
  AGX_TRY(lin160(st, W, tc::TT_PRED0, t.dV, FP, s.u1, FP, W + T.pred0, nullptr, nullptr, nullptr, t.dP, FP, rows, false, false));          // dP_K

  // ---- propagation steps, last to first
  for (int k = K - 1; k >= 0; --k) {
    const float* act = s.P[k + 1];   // ReLU output of this step
    WgradBatch ws(st, t.part);
    AGX_TRY(ws.add(t.dP, act, s.agg[k], FP, FP, rows, F, F, 2 * F, F, gw[AGX_W_PPROP], nullptr));
    mask_rows_kernel<<<(unsigned)((rows * FP + 255) / 256), 256, 0, st>>>(t.dP, act, rows * FP);                              // dP <- d pre_n
    AGX_LAUNCH_CHECK();
    add_rows_kernel<<<(unsigned)((rows * FP + 255) / 256), 256, 0, st>>>(t.dA, t.dP, rows * FP);
    AGX_LAUNCH_CHECK();
    AGX_TRY(lin160(st, W, tc::TT_PP_AGG, t.dP, FP, nullptr, 0, W + T.pp_agg, nullptr, nullptr, nullptr, t.dAgg, FP, rows, false, false));
    if (g->E_cap > 0) {
      effect_bwd_recv_kernel<<<(unsigned)((rows + TA_NODES - 1) / TA_NODES), TA_THREADS, 0, st>>>(
          g->row_ptr, g->send, rows, g->N, g->E_cap, s.C, s.Qr[k], s.Qs[k], fwd_tc, reinterpret_cast<const float4*>(t.dAgg),
          reinterpret_cast<float4*>(t.dC), false, reinterpret_cast<float4*>(t.dQr), nullptr);
      AGX_LAUNCH_CHECK();
      effect_bwd_send_kernel<<<(unsigned)((rows + TA_NODES - 1) / TA_NODES), TA_THREADS, 0, st>>>(
          send_ptr, send_perm, g->recv, rows, g->E_cap, s.C, s.Qr[k], s.Qs[k], fwd_tc, reinterpret_cast<const float4*>(t.dAgg),
          reinterpret_cast<float4*>(t.dQs), nullptr);
      AGX_LAUNCH_CHECK();
      AGX_TRY(ws.add(t.dQr, nullptr, s.P[k], FP, FP, rows, F, F, 3 * F, F, gw[AGX_W_RPROP], nullptr));
      AGX_TRY(ws.add(t.dQs, nullptr, s.P[k], FP, FP, rows, F, F, 3 * F, 2 * F, gw[AGX_W_RPROP], nullptr));
    }
    AGX_TRY(ws.flush());   // before dP is accumulated into and dQr / dQs are rewritten
    if (g->E_cap > 0) {
      // dP_k = d pre_n (residual) + dQr*W_recv + dQs*W_send   (accumulated in place: dP already holds d pre_n)
      AGX_TRY(lin160(st, W, tc::TT_RP_RECV, t.dQr, FP, nullptr, 0, W + T.rp_recv, nullptr, nullptr, nullptr, t.dP, FP, rows, false, true));
      AGX_TRY(lin160(st, W, tc::TT_RP_SEND, t.dQs, FP, nullptr, 0, W + T.rp_send, nullptr, nullptr, nullptr, t.dP, FP, rows, false, true));
    }
  }
  // ---- encoders: d penc = dP_0 + dA * W_enc
  AGX_TRY(wb_hold(wb, t.dA, nullptr, s.penc, FP, FP, rows, F, F, 2 * F, 0, gw[AGX_W_PPROP], gb[AGX_W_PPROP]));
  AGX_TRY(lin160(st, W, tc::TT_PP_ENC, t.dA, FP, nullptr, 0, W + T.pp_enc, nullptr, nullptr, nullptr, t.dP, FP, rows, false, true));
  AGX_TRY(wb_hold(wb, t.dP, s.penc, s.h2, FP, FP, rows, F, F, F, 0, gw[AGX_W_PENC4], gb[AGX_W_PENC4]));
  AGX_TRY(lin160(st, W, tc::TT_PENC4, t.dP, FP, s.penc, FP, W + T.penc4, nullptr, nullptr, nullptr, t.dH2, FP, rows, false, false));
  AGX_TRY(wb_hold(wb, t.dH2, s.h2, s.h1, FP, FP, rows, F, F, F, 0, gw[AGX_W_PENC2], gb[AGX_W_PENC2]));
  AGX_TRY(lin160(st, W, tc::TT_PENC2, t.dH2, FP, s.h2, FP, W + T.penc2, nullptr, nullptr, nullptr, t.dH1, FP, rows, false, false));
  AGX_TRY(wb_hold(wb, t.dH1, s.h1, s.p_in, D_NODE_IN, D_NODE_IN, rows, F, d_node, d_node, 0, gw[AGX_W_PENC0], gb[AGX_W_PENC0]));
  if (g->E_cap > 0) {
    AGX_TRY(wb_hold(wb, t.dC, nullptr, s.renc, FP, FP, E, F, F, 3 * F, 0, gw[AGX_W_RPROP], gb[AGX_W_RPROP], g->row_ptr + rows));
    AGX_TRY(lin160(st, W, tc::TT_RP_REL, t.dC, FP, nullptr, 0, W + T.rp_rel, nullptr, nullptr, nullptr, t.dE, FP, E, false, false));         // dRenc
    AGX_TRY(wb_hold(wb, t.dE, s.renc, s.g2, FP, FP, E, F, F, F, 0, gw[AGX_W_RENC4], gb[AGX_W_RENC4], g->row_ptr + rows));
    AGX_TRY(lin160(st, W, tc::TT_RENC4, t.dE, FP, s.renc, FP, W + T.renc4, nullptr, nullptr, nullptr, t.dG2, FP, E, false, false));
    AGX_TRY(wb_hold(wb, t.dG2, s.g2, s.g1, FP, FP, E, F, F, F, 0, gw[AGX_W_RENC2], gb[AGX_W_RENC2], g->row_ptr + rows));
    AGX_TRY(lin160(st, W, tc::TT_RENC2, t.dG2, FP, s.g2, FP, W + T.renc2, nullptr, nullptr, nullptr, t.dG1, FP, E, false, false));
    AGX_TRY(wb_hold(wb, t.dG1, s.g1, s.rel_in, D_REL_IN, D_REL_IN, E, F, d_rel, d_rel, 0, gw[AGX_W_RENC0], gb[AGX_W_RENC0], g->row_ptr + rows));
    if (d_state)
      AGX_TRY(lin160(st, W, tc::TT_RENC0, t.dG1, FP, s.g1, FP, W + T.renc0, nullptr, nullptr, nullptr, t.dRel, D_REL_IN, E, false, false, D_REL_IN));   // d rel_in
  }
  AGX_TRY(wb.flush());
  if (d_state) {
    if (g->E_cap == 0) AGX_CUDA_OK(cudaMemsetAsync(t.dRel, 0, sizeof(float), st));
    state_bwd_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, st>>>(g->row_ptr, g->E_cap > 0 ? send_ptr : g->row_ptr, send_perm, g->E_cap, t.dRel,
                                                                    d_pred_pos, g->B, g->N, g->n_p, d_state);
    AGX_LAUNCH_CHECK();
  }
  return AGX_OK;
}

}  // extern "C"
