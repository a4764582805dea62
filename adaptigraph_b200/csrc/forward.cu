// DynamicsPredictor.forward (reference dynamics/gnn/model.py:129-313) on CSR relations, and the
// autoregressive rollout loop (planning/forward_dynamics.py:156-197), as fused sm_100a kernels.
//
// The reference evaluates every gather / scatter as a dense bmm against B x n_rel x N one-hots and
// the relation propagator as a 450->150 product per relation per propagation step.  Here:
//
//   * relations are CSR by receiver (row_ptr / send / recv), nothing is O(n_rel * N);
//   * the two propagators are split by operand (exact algebra, model.py:288-289, :299-301):
//         relation_propagator([renc | P[recv] | P[send]]) = relu(C_e + Qr[recv] + Qs[send])
//             C_e = W_rel * renc + b   (constant over psteps, one product per relation in total)
//             Qr  = P * W_recv^T,  Qs = P * W_send^T   (per node, per pstep)
//         particle_propagator([penc | agg], res=P) = relu(A_n + W_agg * agg + P),  A_n = W_enc * penc + b
//   * kernels (AGX_PREC_FP32 path; each is one fused chain over a 128-row tile):
//         node_encoder    inputs -> nfeat record, penc (3 layers) -> P, A_n, Qr, Qs
//         edge_encoder    gather nfeat[recv], nfeat[send] -> 17 relation inputs -> renc (3 layers) -> C_e
//         edge_aggregate  agg[n] = sum_{e in row n} relu(C_e + Qr[n] + Qs[send[e]])   (HBM/L2-bound)
//         node_update     P <- relu(A_n + W_agg*agg + P); then Qr, Qs for the next pstep, or on the
//                         last pstep the motion head (model.py:306-309) and pred_pos = pos + clamp(motion)
//         rollout_advance tool kinematics + history shift between rollout steps
#include "common.cuh"
#include "tc_forward.cuh"
#include "mlp_simt.cuh"

namespace agx {

int graph_build_impl(const float* pos, int64_t pos_stride_b, const uint8_t* mask, const uint8_t* tool_mask,
                     const float* thr2, int B, int N, int topk, int cta, int sem, int32_t* row_ptr, int32_t* send,
                     int32_t* recv, int64_t cap, int32_t* n_edges, int32_t* status, void* workspace,
                     size_t workspace_bytes, cudaStream_t st);

constexpr float MOTION_CLAMP = 100.f;  // model.py:85

// tensor-core path (tc_forward.cu): tc_forward.cuh
size_t tc_blob_bytes(size_t base_bytes);
size_t train_blob_bytes();
int train_pack(const AgxModelDims* dims, const AgxWeights* raw, void* packed, cudaStream_t st);
int tc_pack(const AgxModelDims* dims, const AgxWeights* raw, void* packed, size_t base_bytes, cudaStream_t st);

// ------------------------------------------------------------------------------------ weight packing
__global__ void pack_mat_kernel(const float* __restrict__ W, int ld, int col0, int K, int F, int Kpad, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Kpad * FP) return;
  const int k = i / FP, n = i - k * FP;
  dst[i] = (k < K && n < F) ? W[(size_t)n * ld + col0 + k] : 0.f;
}
__global__ void pack_vec_kernel(const float* __restrict__ b, int n_valid, int n_pad, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pad) dst[i] = i < n_valid ? b[i] : 0.f;
}
__global__ void pack_rows_kernel(const float* __restrict__ W, int n_out, int F, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out * FP) return;
  const int o = i / FP, k = i - o * FP;
  dst[i] = k < F ? W[(size_t)o * F + k] : 0.f;
}

// ------------------------------------------------------------------------------------ node encoder
struct NodeEncArgs {
  const float* state; const float* attrs; const float* action; const float* p_instance; const float* physics;
  int B, N, n_p;
  const float* wts;  // packed blob
  PackedLayout L;
  float* nfeat; float* P; float* A; float* Qr; float* Qs;
};

template <bool TO_SMEM_AND_GLOBAL>
__device__ __forceinline__ void epilogue_relu_smem_global(float* Xs, Acc& acc, const float* __restrict__ bias, float* __restrict__ out,
                                                          int64_t row0, int64_t n_rows, int tid) {
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
    if (TO_SMEM_AND_GLOBAL) {
      const int64_t gr = row0 + 8 * tr + r;
      if (gr < n_rows) {
        float* g = out + gr * FP;
        *reinterpret_cast<float4*>(g + 4 * tc) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(g + 64 + 4 * tc) = make_float4(o[4], o[5], o[6], o[7]);
        *reinterpret_cast<float2*>(g + 128 + 2 * tc) = make_float2(o[8], o[9]);
      }
    }
  }
}

__global__ void __launch_bounds__(MLP_THREADS, 2) node_encoder_kernel(const NodeEncArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* Xs = smem;
  float* Ws = smem + TM * LDX;
  const int tid = threadIdx.x;
  const int64_t rows = (int64_t)a.B * a.N;
  const int n_tiles = (int)((rows + TM - 1) / TM);
  const float* W = a.wts;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = (int64_t)tile * TM;
    __syncthreads();
    if (tid < TM) {
      const int64_t r = row0 + tid;
      float in[D_NODE_IN] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (r < rows) {
        const int b = (int)(r / a.N), n = (int)(r - (int64_t)b * a.N);
        float s[H_FIX][3];
#pragma unroll
        for (int h = 0; h < H_FIX; ++h) {
          const float* p = a.state + (((size_t)b * H_FIX + h) * a.N + n) * 3;
          s[h][0] = p[0]; s[h][1] = p[1]; s[h][2] = p[2];
        }
        const float a0 = a.attrs[r * 2 + 0], a1 = a.attrs[r * 2 + 1];
        const float grp = n < a.n_p ? a.p_instance[(size_t)b * a.n_p + n] : 0.f;
        float4* nf = reinterpret_cast<float4*>(a.nfeat + r * NFEAT);
        // history record (model.py:155-165): [s1-s0, s2-s1, s3-s2, s3]
        nf[0] = make_float4(s[1][0] - s[0][0], s[1][1] - s[0][1], s[1][2] - s[0][2], s[2][0] - s[1][0]);
        nf[1] = make_float4(s[2][1] - s[1][1], s[2][2] - s[1][2], s[3][0] - s[2][0], s[3][1] - s[2][1]);
        nf[2] = make_float4(s[3][2] - s[2][2], s[3][0], s[3][1], s[3][2]);
        nf[3] = make_float4(a0, a1, grp, 0.f);
        in[0] = a0; in[1] = a1;
        in[2] = n < a.n_p ? a.physics[b] : 0.f;           // model.py:186-189
        in[3] = a.action[r * 3 + 0]; in[4] = a.action[r * 3 + 1]; in[5] = a.action[r * 3 + 2];
      }
      float4* x = reinterpret_cast<float4*>(Xs + (size_t)tid * LDX);
      x[0] = make_float4(in[0], in[1], in[2], in[3]);
      x[1] = make_float4(in[4], in[5], in[6], in[7]);
    }
    Acc acc;
    tile_gemm<D_NODE_IN>(Xs, W + a.L.penc0_w, Ws, acc, tid);
    epilogue_bias_relu_to_smem(Xs, acc, W + a.L.penc0_b, tid);
    tile_gemm<FP>(Xs, W + a.L.penc2_w, Ws, acc, tid);
    epilogue_bias_relu_to_smem(Xs, acc, W + a.L.penc2_b, tid);
    tile_gemm<FP>(Xs, W + a.L.penc4_w, Ws, acc, tid);
    epilogue_relu_smem_global<true>(Xs, acc, W + a.L.penc4_b, a.P, row0, rows, tid);   // particle_effect_0 = particle_encode (:269)
    tile_gemm<FP>(Xs, W + a.L.pp_enc_w, Ws, acc, tid);
    add_bias(acc, W + a.L.pp_b, tid);
    store_rows(a.A, row0, rows, acc, tid);
    tile_gemm<FP>(Xs, W + a.L.rp_recv_w, Ws, acc, tid);
    store_rows(a.Qr, row0, rows, acc, tid);
    tile_gemm<FP>(Xs, W + a.L.rp_send_w, Ws, acc, tid);
    store_rows(a.Qs, row0, rows, acc, tid);
  }
}

// ------------------------------------------------------------------------------------ edge encoder
struct EdgeEncArgs {
  const int32_t* row_ptr; const int32_t* send; const int32_t* recv;
  int64_t rows; int N; int64_t E_cap;
  const float* nfeat;
  const float* wts; PackedLayout L;
  float* C;
};

__global__ void __launch_bounds__(MLP_THREADS, 2) edge_encoder_kernel(const EdgeEncArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* Xs = smem;
  float* Ws = smem + TM * LDX;
  const int tid = threadIdx.x;
  const int64_t E = min((int64_t)a.row_ptr[a.rows], a.E_cap);
  const int n_tiles = (int)((E + TM - 1) / TM);
  const float* W = a.wts;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t e0 = (int64_t)tile * TM;
    __syncthreads();
    if (tid < TM) {
      const int64_t e = e0 + tid;
      float in[D_REL_IN];
#pragma unroll
      for (int i = 0; i < D_REL_IN; ++i) in[i] = 0.f;
      if (e < E) {
        const int r = a.recv[e];
        const int s = (r / a.N) * a.N + a.send[e];
        const float4* fr = reinterpret_cast<const float4*>(a.nfeat + (size_t)r * NFEAT);
        const float4* fs = reinterpret_cast<const float4*>(a.nfeat + (size_t)s * NFEAT);
        const float4 r0 = fr[0], r1 = fr[1], r2 = fr[2], r3 = fr[3];
        const float4 s0 = fs[0], s1 = fs[1], s2 = fs[2], s3 = fs[3];
        // model.py:224-253: [attr_r, attr_s, sum|group_r - group_s|, hist_r - hist_s]
        in[0] = r3.x; in[1] = r3.y; in[2] = s3.x; in[3] = s3.y;
        in[4] = fabsf(r3.z - s3.z);
        in[5] = r0.x - s0.x; in[6] = r0.y - s0.y; in[7] = r0.z - s0.z; in[8] = r0.w - s0.w;
        in[9] = r1.x - s1.x; in[10] = r1.y - s1.y; in[11] = r1.z - s1.z; in[12] = r1.w - s1.w;
        in[13] = r2.x - s2.x; in[14] = r2.y - s2.y; in[15] = r2.z - s2.z; in[16] = r2.w - s2.w;
      }
      float4* x = reinterpret_cast<float4*>(Xs + (size_t)tid * LDX);
#pragma unroll
      for (int i = 0; i < D_REL_IN / 4; ++i) x[i] = make_float4(in[4 * i], in[4 * i + 1], in[4 * i + 2], in[4 * i + 3]);
    }
    Acc acc;
    tile_gemm<D_REL_IN>(Xs, W + a.L.renc0_w, Ws, acc, tid);
    epilogue_bias_relu_to_smem(Xs, acc, W + a.L.renc0_b, tid);
    tile_gemm<FP>(Xs, W + a.L.renc2_w, Ws, acc, tid);
    epilogue_bias_relu_to_smem(Xs, acc, W + a.L.renc2_b, tid);
    tile_gemm<FP>(Xs, W + a.L.renc4_w, Ws, acc, tid);
    epilogue_bias_relu_to_smem(Xs, acc, W + a.L.renc4_b, tid);
    tile_gemm<FP>(Xs, W + a.L.rp_rel_w, Ws, acc, tid);
    add_bias(acc, W + a.L.rp_b, tid);
    store_rows(a.C, e0, E, acc, tid);
  }
}

// ------------------------------------------------------------------------------------ edge aggregate
constexpr int AGG_NODES = 8;
constexpr int AGG_THREADS = AGG_NODES * (FP / 4);  // 320

__device__ __forceinline__ float4 relu_add3(const float4 c, const float4 qr, const float4 qs) {
  return make_float4(fmaxf((c.x + qr.x) + qs.x, 0.f), fmaxf((c.y + qr.y) + qs.y, 0.f),
                     fmaxf((c.z + qr.z) + qs.z, 0.f), fmaxf((c.w + qr.w) + qs.w, 0.f));
}

__global__ void __launch_bounds__(AGG_THREADS) edge_aggregate_kernel(
    const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ send, int64_t rows, int N, int64_t E_cap,
    const float4* __restrict__ C, const float4* __restrict__ Qr, const float4* __restrict__ Qs, float4* __restrict__ agg) {
  const int slot = threadIdx.x / (FP / 4), j = threadIdx.x - slot * (FP / 4);
  const int64_t r = (int64_t)blockIdx.x * AGG_NODES + slot;
  if (r >= rows) return;
  const int64_t beg = row_ptr[r];
  const int64_t end = min((int64_t)row_ptr[r + 1], E_cap);
  const int64_t gb = (r / N) * N;
  const float4 qr = Qr[r * (FP / 4) + j];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int64_t e = beg;
  for (; e + 1 < end; e += 2) {
    const int64_t s0 = gb + send[e], s1 = gb + send[e + 1];
    const float4 c0 = C[e * (FP / 4) + j], c1 = C[(e + 1) * (FP / 4) + j];
    const float4 q0 = Qs[s0 * (FP / 4) + j], q1 = Qs[s1 * (FP / 4) + j];
    const float4 v0 = relu_add3(c0, qr, q0), v1 = relu_add3(c1, qr, q1);
    acc.x += v0.x; acc.y += v0.y; acc.z += v0.z; acc.w += v0.w;
    acc.x += v1.x; acc.y += v1.y; acc.z += v1.z; acc.w += v1.w;
  }
  if (e < end) {
    const int64_t s0 = gb + send[e];
    const float4 v0 = relu_add3(C[e * (FP / 4) + j], qr, Qs[s0 * (FP / 4) + j]);
    acc.x += v0.x; acc.y += v0.y; acc.z += v0.z; acc.w += v0.w;
  }
  agg[r * (FP / 4) + j] = acc;
}

// ------------------------------------------------------------------------------------ node update / head
struct NodeUpdArgs {
  int B, N, n_p;
  const float* agg; const float* A; float* P; float* Qr; float* Qs;
  const float* wts; PackedLayout L;
  // head (last pstep only)
  const float* state; float* pred_pos; int64_t pos_stride_b; float* pred_motion;
};

template <bool LAST>
__global__ void __launch_bounds__(MLP_THREADS, 2) node_update_kernel(const NodeUpdArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* Xs = smem;
  float* Ws = smem + TM * LDX;
  const int tid = threadIdx.x, tr = tid >> 4, tc = tid & 15;
  const int64_t rows = (int64_t)a.B * a.N;
  const int n_tiles = (int)((rows + TM - 1) / TM);
  const float* W = a.wts;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = (int64_t)tile * TM;
    __syncthreads();
    for (int i = tid; i < TM * (FP / 4); i += MLP_THREADS) {
      const int row = i / (FP / 4), c4 = i - row * (FP / 4);
      const int64_t gr = row0 + row;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gr < rows) v = reinterpret_cast<const float4*>(a.agg)[gr * (FP / 4) + c4];
      *reinterpret_cast<float4*>(Xs + (size_t)row * LDX + 4 * c4) = v;
    }
    Acc acc;
    tile_gemm<FP>(Xs, W + a.L.pp_agg_w, Ws, acc, tid);
    // P <- relu((W_agg*agg + A_n) + P)   (model.py:36-40, :299-301)
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int64_t gr = row0 + 8 * tr + r;
      float o[10];
      if (gr < rows) {
        const float* Ar = a.A + gr * FP;
        float* Pr = a.P + gr * FP;
        const float4 a0 = *reinterpret_cast<const float4*>(Ar + 4 * tc), a1 = *reinterpret_cast<const float4*>(Ar + 64 + 4 * tc);
        const float2 a2 = *reinterpret_cast<const float2*>(Ar + 128 + 2 * tc);
        const float4 p0 = *reinterpret_cast<const float4*>(Pr + 4 * tc), p1 = *reinterpret_cast<const float4*>(Pr + 64 + 4 * tc);
        const float2 p2 = *reinterpret_cast<const float2*>(Pr + 128 + 2 * tc);
        o[0] = fmaxf((acc.v[r][0] + a0.x) + p0.x, 0.f); o[1] = fmaxf((acc.v[r][1] + a0.y) + p0.y, 0.f);
        o[2] = fmaxf((acc.v[r][2] + a0.z) + p0.z, 0.f); o[3] = fmaxf((acc.v[r][3] + a0.w) + p0.w, 0.f);
        o[4] = fmaxf((acc.v[r][4] + a1.x) + p1.x, 0.f); o[5] = fmaxf((acc.v[r][5] + a1.y) + p1.y, 0.f);
        o[6] = fmaxf((acc.v[r][6] + a1.z) + p1.z, 0.f); o[7] = fmaxf((acc.v[r][7] + a1.w) + p1.w, 0.f);
        o[8] = fmaxf((acc.v[r][8] + a2.x) + p2.x, 0.f); o[9] = fmaxf((acc.v[r][9] + a2.y) + p2.y, 0.f);
        if (!LAST) {
          *reinterpret_cast<float4*>(Pr + 4 * tc) = make_float4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<float4*>(Pr + 64 + 4 * tc) = make_float4(o[4], o[5], o[6], o[7]);
          *reinterpret_cast<float2*>(Pr + 128 + 2 * tc) = make_float2(o[8], o[9]);
        }
      } else {
#pragma unroll
        for (int c = 0; c < 10; ++c) o[c] = 0.f;
      }
      float* row = Xs + (size_t)(8 * tr + r) * LDX;
      *reinterpret_cast<float4*>(row + 4 * tc) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4*>(row + 64 + 4 * tc) = make_float4(o[4], o[5], o[6], o[7]);
      *reinterpret_cast<float2*>(row + 128 + 2 * tc) = make_float2(o[8], o[9]);
    }
    if (!LAST) {
      tile_gemm<FP>(Xs, W + a.L.rp_recv_w, Ws, acc, tid);
      store_rows(a.Qr, row0, rows, acc, tid);
      tile_gemm<FP>(Xs, W + a.L.rp_send_w, Ws, acc, tid);
      store_rows(a.Qs, row0, rows, acc, tid);
    } else {
      tile_gemm<FP>(Xs, W + a.L.pred0_w, Ws, acc, tid);
      epilogue_bias_relu_to_smem(Xs, acc, W + a.L.pred0_b, tid);
      tile_gemm<FP>(Xs, W + a.L.pred1_w, Ws, acc, tid);
      epilogue_bias_relu_to_smem(Xs, acc, W + a.L.pred1_b, tid);
      __syncthreads();
      if (tid < TM) {
        const int64_t gr = row0 + tid;
        if (gr < rows) {
          const int b = (int)(gr / a.N), n = (int)(gr - (int64_t)b * a.N);
          if (n < a.n_p) {
            const float* x = Xs + (size_t)tid * LDX;
            const float* w2 = W + a.L.pred2_w;
            float m0 = 0.f, m1 = 0.f, m2 = 0.f;
#pragma unroll 4
            for (int k = 0; k < FP; k += 4) {
              const float4 xv = *reinterpret_cast<const float4*>(x + k);
              const float4 u0 = *reinterpret_cast<const float4*>(w2 + k);
              const float4 u1 = *reinterpret_cast<const float4*>(w2 + FP + k);
              const float4 u2 = *reinterpret_cast<const float4*>(w2 + 2 * FP + k);
              m0 = fmaf(xv.x, u0.x, m0); m0 = fmaf(xv.y, u0.y, m0); m0 = fmaf(xv.z, u0.z, m0); m0 = fmaf(xv.w, u0.w, m0);
              m1 = fmaf(xv.x, u1.x, m1); m1 = fmaf(xv.y, u1.y, m1); m1 = fmaf(xv.z, u1.z, m1); m1 = fmaf(xv.w, u1.w, m1);
              m2 = fmaf(xv.x, u2.x, m2); m2 = fmaf(xv.y, u2.y, m2); m2 = fmaf(xv.z, u2.z, m2); m2 = fmaf(xv.w, u2.w, m2);
            }
            const float* b2 = W + a.L.pred2_b;
            m0 += b2[0]; m1 += b2[1]; m2 += b2[2];
            float* mo = a.pred_motion + ((size_t)b * a.n_p + n) * 3;
            mo[0] = m0; mo[1] = m1; mo[2] = m2;
            const float* cur = a.state + (((size_t)b * H_FIX + (H_FIX - 1)) * a.N + n) * 3;
            float* po = a.pred_pos + (size_t)b * a.pos_stride_b + (size_t)n * 3;
            po[0] = cur[0] + fminf(fmaxf(m0, -MOTION_CLAMP), MOTION_CLAMP);   // model.py:309
            po[1] = cur[1] + fminf(fmaxf(m1, -MOTION_CLAMP), MOTION_CLAMP);
            po[2] = cur[2] + fminf(fmaxf(m2, -MOTION_CLAMP), MOTION_CLAMP);
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------ rollout advance
// forward_dynamics.py:163-176: tools move by their action delta and take y from the predicted
// object heights; the history drops its oldest frame and appends [pred ; tools].
// With nfeat != nullptr the kernel also writes the NEXT step's history records (model.py:155-165 on the advanced history; the
// tensor-core path's relation encoder gathers them), so that a later rollout step needs no particle-side kernel before its
// relation encoder: the particle encoder's products are reused (agx_rollout).
__global__ void __launch_bounds__(256) rollout_advance_kernel(float* __restrict__ state, const float* __restrict__ action,
                                                               const uint8_t* __restrict__ mask, const float* __restrict__ pred,
                                                               int64_t pred_stride_b, int N, int n_p, int y_mode, float raise,
                                                               float* __restrict__ nfeat, const float* __restrict__ attrs,
                                                               const float* __restrict__ p_instance) {
  __shared__ float red_a[8], red_b[8];
  __shared__ float y_s;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_wait();
  pdl_trigger();
  const float* pb = pred + (size_t)b * pred_stride_b;
  float va = (y_mode == AGX_Y_MIN) ? __int_as_float(0x7f800000) : 0.f, vb = 0.f;
  for (int n = tid; n < n_p; n += 256) {
    const float y = pb[n * 3 + 1];
    if (y_mode == AGX_Y_MIN) va = fminf(va, y);
    else { const float m = mask[(size_t)b * N + n] ? 1.f : 0.f; va += y * m; vb += m; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ua = __shfl_xor_sync(0xffffffffu, va, o), ub = __shfl_xor_sync(0xffffffffu, vb, o);
    va = (y_mode == AGX_Y_MIN) ? fminf(va, ua) : va + ua;
    vb += ub;
  }
  if (lane == 0) { red_a[warp] = va; red_b[warp] = vb; }
  __syncthreads();
  if (tid == 0) {
    float ta = red_a[0], tb = red_b[0];
    for (int w = 1; w < 8; ++w) { ta = (y_mode == AGX_Y_MIN) ? fminf(ta, red_a[w]) : ta + red_a[w]; tb += red_b[w]; }
    y_s = ((y_mode == AGX_Y_MIN) ? ta : ta / tb) + raise;
  }
  __syncthreads();
  const float y_tool = y_s;
  // grid = (graphs, chunks of 256 particles): every CTA repeats the (cheap) reduction over the graph's predictions above and then
  // advances its own chunk, so that small batches are not left to one CTA per graph
  for (int n = blockIdx.y * 256 + tid; n < min(N, (int)(blockIdx.y + 1) * 256); n += 256) {
    float s[H_FIX][3];
#pragma unroll
    for (int h = 0; h < H_FIX; ++h) {
      const float* p = state + (((size_t)b * H_FIX + h) * N + n) * 3;
      s[h][0] = p[0]; s[h][1] = p[1]; s[h][2] = p[2];
    }
    float nx, ny, nz;
    if (n < n_p) { nx = pb[n * 3]; ny = pb[n * 3 + 1]; nz = pb[n * 3 + 2]; }
    else {
      const float* ac = action + ((size_t)b * N + n) * 3;
      nx = s[H_FIX - 1][0] + ac[0]; ny = y_tool; nz = s[H_FIX - 1][2] + ac[2];
    }
#pragma unroll
    for (int h = 0; h < H_FIX - 1; ++h) {
      float* p = state + (((size_t)b * H_FIX + h) * N + n) * 3;
      p[0] = s[h + 1][0]; p[1] = s[h + 1][1]; p[2] = s[h + 1][2];
    }
    float* p = state + (((size_t)b * H_FIX + (H_FIX - 1)) * N + n) * 3;
    p[0] = nx; p[1] = ny; p[2] = nz;
    if (nfeat) {   // the shifted history is s[1], s[2], s[3], (nx, ny, nz): same record as tc_forward.cu's node_inputs
      const size_t r = (size_t)b * N + n;
      float4* nf = reinterpret_cast<float4*>(nfeat + r * NFEAT);
      nf[0] = make_float4(s[2][0] - s[1][0], s[2][1] - s[1][1], s[2][2] - s[1][2], s[3][0] - s[2][0]);
      nf[1] = make_float4(s[3][1] - s[2][1], s[3][2] - s[2][2], nx - s[3][0], ny - s[3][1]);
      nf[2] = make_float4(nz - s[3][2], nx, ny, nz);
      nf[3] = make_float4(attrs[r * 2 + 0], attrs[r * 2 + 1], n < n_p ? p_instance[(size_t)b * n_p + n] : 0.f, 0.f);
    }
  }
}

// ------------------------------------------------------------------------------------ host drivers
struct FwdWs {
  float* nfeat; float* P; float* A; float* Qr; float* Qs; float* agg; float* C; float* rowmaxP; float* rowmaxA;
  int32_t* agg_exp; float* agg_max;
  float* P0; float* Qr0; float* Qs0; float* rowmaxP0;   // tensor-core path: the particle encoder's products, kept across rollout steps
  float* S0;
};
static size_t fwd_ws_carve(void* base, int64_t rows, int64_t E_cap, FwdWs* out) {
  Carver c(base);
  FwdWs w;
  w.nfeat = c.take<float>((size_t)rows * NFEAT);
  // padded to whole 128-row tiles: the tensor-core path keeps these matrices tile-blocked (tc_chain.cuh: blk_off)
  const size_t rows_pad = (size_t)((rows + 127) / 128 * 128), e_pad = (size_t)(((E_cap > 0 ? E_cap : 1) + 127) / 128 * 128);
  w.P = c.take<float>(rows_pad * FP);
  w.A = c.take<float>(rows_pad * FP);
  w.Qr = c.take<float>(rows_pad * FP);
  w.Qs = c.take<float>(rows_pad * FP);
  w.agg = c.take<float>(rows_pad * FP);
  w.C = c.take<float>(e_pad * FP);
  w.rowmaxP = c.take<float>((size_t)rows);
  w.rowmaxA = c.take<float>((size_t)rows);
  w.agg_exp = c.take<int32_t>((size_t)rows);
  w.agg_max = c.take<float>((size_t)rows);
  w.P0 = c.take<float>(rows_pad * FP);
  w.Qr0 = c.take<float>(rows_pad * FP);
  w.Qs0 = c.take<float>(rows_pad * FP);
  w.rowmaxP0 = c.take<float>((size_t)rows);
  w.S0 = c.take<float>(rows_pad * FP);
  if (out) *out = w;
  return align_up(c.off, 256);
}

static int check_dims(const AgxModelDims* d) {
  AGX_REQUIRE(d, AGX_ERR_ARG, "dims is null");
  AGX_REQUIRE(d->F >= 8 && d->F <= FP, AGX_ERR_ARG, "F=%d unsupported (need 8..%d)", d->F, FP);
  AGX_REQUIRE(d->n_his == H_FIX, AGX_ERR_ARG, "n_his=%d unsupported (kernels are specialised for %d)", d->n_his, H_FIX);
  AGX_REQUIRE(d->d_attr == 2 && d->d_phys == 1 && d->d_act == 3, AGX_ERR_ARG,
              "attr_dim/physics/action dims (%d,%d,%d) unsupported (need 2,1,3)", d->d_attr, d->d_phys, d->d_act);
  AGX_REQUIRE(d->pstep >= 1, AGX_ERR_ARG, "pstep=%d must be >= 1", d->pstep);
  return AGX_OK;
}

static int ensure_smem_attrs() {
  static thread_local DeviceOnce once;
  if (!once.need()) return AGX_OK;
  AGX_CUDA_OK(cudaFuncSetAttribute(node_encoder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MLP_SMEM_BYTES));
  AGX_CUDA_OK(cudaFuncSetAttribute(edge_encoder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MLP_SMEM_BYTES));
  AGX_CUDA_OK(cudaFuncSetAttribute(node_update_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MLP_SMEM_BYTES));
  AGX_CUDA_OK(cudaFuncSetAttribute(node_update_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MLP_SMEM_BYTES));
  return AGX_OK;
}

// reuse_node_products: the workspace still holds the particle encoder's products of an earlier call on the same attrs / physics /
// action (a later step of a rollout): only the history records are refreshed (tensor-core path; the fp32 path recomputes)
static int forward_impl(const AgxModelDims* dims, const float* wts, const AgxGraphIn* g, float* pred_pos,
                        int64_t pos_stride_b, float* pred_motion, int precision, void* workspace, size_t workspace_bytes,
                        cudaStream_t st, cudaEvent_t after_encoders = nullptr, bool reuse_node_products = false,
                        bool nfeat_ready = false) {
  AGX_REQUIRE(precision == AGX_PREC_FP32 || precision == AGX_PREC_TC_F16X3 || precision == AGX_PREC_TC_MIXED, AGX_ERR_ARG,
              "unknown precision %d", precision);
  const int64_t rows = (int64_t)g->B * g->N;
  FwdWs ws;
  const size_t need = fwd_ws_carve(workspace, rows, g->E_cap, &ws);
  AGX_REQUIRE(workspace && workspace_bytes >= need, AGX_ERR_CAPACITY, "forward: workspace %zu < %zu bytes", workspace_bytes, need);
  if (int rc = ensure_smem_attrs()) return rc;
  const PackedLayout L = packed_layout();
  if (precision != AGX_PREC_FP32) {
    // same stages, dense layers on the tcgen05 tensor cores (tc_forward.cu); MIXED: relation chain at 2 MMAs per K step, C as C16
    const bool mixed = precision == AGX_PREC_TC_MIXED;
    const size_t base = L.total * sizeof(float);
    TcFwdBuffers tb{ws.nfeat, ws.P, ws.A, ws.Qr, ws.Qs, ws.agg, ws.C, ws.rowmaxP, ws.rowmaxA, ws.agg_exp, ws.agg_max,
                    ws.P0, ws.Qr0, ws.Qs0, ws.rowmaxP0, ws.S0};
    tb.agg_f32 = mixed;
    if (reuse_node_products) {
      if (!nfeat_ready)                          // (agx_rollout's advance kernel has normally written the records already)
        if (int rc = tc_nfeat(g, tb, st)) return rc;
    } else {
      if (int rc = tc_node_encoder(g, wts, L, base, tb, st)) return rc;
    }
    if (g->E_cap > 0)
      if (int rc = tc_edge_encoder(g, wts, L, base, tb, mixed, st)) return rc;
    if (after_encoders) AGX_CUDA_OK(cudaEventRecord(after_encoders, st));   // the tensor-bound phase of this step is enqueued
    for (int k = 0; k < dims->pstep; ++k) {
      if (int rc = tc_edge_aggregate(g, tb, mixed, k == 0, st)) return rc;
      if (int rc = tc_node_update(g, wts, L, base, tb, k == 0, k + 1 == dims->pstep, pred_pos, pos_stride_b, pred_motion, st)) return rc;
    }
    return AGX_OK;
  }
  const int grid_cap = 2 * num_sms();
  const int node_tiles = (int)((rows + TM - 1) / TM);
  const int node_grid = node_tiles < grid_cap ? node_tiles : grid_cap;

  NodeEncArgs na{g->state, g->attrs, g->action, g->p_instance, g->physics, g->B, g->N, g->n_p, wts, L,
                 ws.nfeat, ws.P, ws.A, ws.Qr, ws.Qs};
  { ProfScope ps(AGX_KIND_NODE_ENCODER, st);
    node_encoder_kernel<<<node_grid, MLP_THREADS, MLP_SMEM_BYTES, st>>>(na); }
  AGX_LAUNCH_CHECK();

  if (g->E_cap > 0) {
    const int64_t edge_tiles = (g->E_cap + TM - 1) / TM;
    const int edge_grid = (int)(edge_tiles < grid_cap ? edge_tiles : grid_cap);
    EdgeEncArgs ea{g->row_ptr, g->send, g->recv, rows, g->N, g->E_cap, ws.nfeat, wts, L, ws.C};
    { ProfScope ps(AGX_KIND_EDGE_ENCODER, st);
      edge_encoder_kernel<<<edge_grid, MLP_THREADS, MLP_SMEM_BYTES, st>>>(ea); }
    AGX_LAUNCH_CHECK();
  }

  NodeUpdArgs ua{g->B, g->N, g->n_p, ws.agg, ws.A, ws.P, ws.Qr, ws.Qs, wts, L, g->state, pred_pos, pos_stride_b, pred_motion};
  for (int k = 0; k < dims->pstep; ++k) {
    { ProfScope ps(AGX_KIND_EDGE_AGGREGATE, st);
      edge_aggregate_kernel<<<(unsigned)((rows + AGG_NODES - 1) / AGG_NODES), AGG_THREADS, 0, st>>>(
          g->row_ptr, g->send, rows, g->N, g->E_cap, reinterpret_cast<const float4*>(ws.C),
          reinterpret_cast<const float4*>(ws.Qr), reinterpret_cast<const float4*>(ws.Qs), reinterpret_cast<float4*>(ws.agg)); }
    AGX_LAUNCH_CHECK();
    if (k + 1 < dims->pstep) {
      ProfScope ps(AGX_KIND_NODE_UPDATE, st);
      node_update_kernel<false><<<node_grid, MLP_THREADS, MLP_SMEM_BYTES, st>>>(ua);
    } else {
      ProfScope ps(AGX_KIND_NODE_HEAD, st);
      node_update_kernel<true><<<node_grid, MLP_THREADS, MLP_SMEM_BYTES, st>>>(ua);
    }
    AGX_LAUNCH_CHECK();
  }
  return AGX_OK;
}

struct RolloutWs {
  int32_t* row_ptr; int32_t* send; int32_t* recv; int32_t* n_edges; float* motion; void* graph_ws; size_t graph_ws_bytes;
  void* fwd_ws; size_t fwd_ws_bytes;
};
static size_t rollout_ws_carve(void* base, int B, int N, int64_t E_cap, int topk, RolloutWs* out) {
  Carver c(base);
  const int64_t rows = (int64_t)B * N;
  RolloutWs w;
  w.row_ptr = c.take<int32_t>(rows + 1);
  w.send = c.take<int32_t>(E_cap);
  w.recv = c.take<int32_t>(E_cap);
  w.n_edges = c.take<int32_t>(B);
  w.motion = c.take<float>((size_t)rows * 3);
  w.graph_ws_bytes = agx_graph_workspace_bytes(B, N, topk);
  w.graph_ws = c.take<char>(w.graph_ws_bytes);
  w.fwd_ws_bytes = fwd_ws_carve(nullptr, rows, E_cap, nullptr);
  w.fwd_ws = c.take<char>(w.fwd_ws_bytes);
  if (out) *out = w;
  return align_up(c.off, 256);
}

// ---- side-by-side half batches (agx_rollout)
struct SplitPlan { int B0; int64_t E0, E1; int sms; };
struct SplitStreams { cudaStream_t side; cudaEvent_t fork, offset, join; };
static SplitStreams* split_streams() {   // per host thread, created on first use, never destroyed (process lifetime)
  static thread_local SplitStreams ss{};
  static thread_local DeviceOnce once;
  if (once.need()) {
    if (cudaStreamCreateWithFlags(&ss.side, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ss.offset, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ss.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  }
  return &ss;
}
// Opt-in (AGX_ROLLOUT_SPLIT=1), tensor-core path only.  Measured on cloth-2k x 128 (B200): 29.3 ms per 10-step rollout against
// 27.9 ms unsplit — run alone on 74 SMs a half batch needs 0.241 / 0.647 ms for edge_aggregate / edge_encoder, but side by side each
// half's kernels take as long as the full-batch kernels on the whole GPU (0.315 / 0.72 ms): two thirds of a step is HBM bound, so the
// two halves mostly compete for HBM instead of interleaving with each other's tensor-bound phase.  Kept for workloads with a
// larger tensor-bound share (more propagation-free encoder work per relation); results are bit-identical either way.
static bool rollout_split_plan(int B, int N, int64_t E_cap, int topk, int precision, SplitPlan* p) {
  (void)topk; (void)N;
  const char* e = getenv("AGX_ROLLOUT_SPLIT");
  if (!e || atoi(e) != 1) return false;
  if (precision == AGX_PREC_FP32 || B < 2) return false;
  const int sms = num_sms() / 2;
  p->B0 = B / 2;
  p->E0 = E_cap / B * p->B0;          // E_cap is per-graph capacity x B for every caller of the Python layer
  p->E1 = E_cap - p->E0;
  p->sms = sms;
  return sms > 0;
}

}  // namespace agx

extern "C" {

size_t agx_packed_weights_bytes(const AgxModelDims* dims) {
  (void)dims;
  return agx::train_blob_bytes();   // fp32 k-major section + tensor-core fp16 images + plain copies for the dgrad products
}

int agx_pack_weights(const AgxModelDims* dims, const AgxWeights* raw, void* packed, agx_stream_t stream) {
  using namespace agx;
  if (int rc = check_dims(dims)) return rc;
  AGX_REQUIRE(raw && packed, AGX_ERR_ARG, "pack_weights: null pointer argument");
  for (int i = 0; i < AGX_NUM_LAYERS; ++i)
    AGX_REQUIRE(raw->weight[i] && raw->bias[i], AGX_ERR_ARG, "pack_weights: layer %d has a null pointer", i);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const PackedLayout L = packed_layout();
  float* out = static_cast<float*>(packed);
  const int F = dims->F;
  const int d_node = dims->d_attr + dims->d_phys + dims->d_act;                 // model.py:96-101 (state_dim 0)
  const int d_rel = 2 * dims->d_attr + 1 + 3 * dims->n_his;                     // model.py:109-113
  AGX_REQUIRE(d_node <= D_NODE_IN && d_rel <= D_REL_IN, AGX_ERR_ARG, "input dims (%d,%d) exceed the padded sizes", d_node, d_rel);
  auto mat = [&](int layer, int ld, int col0, int K, int Kpad, size_t off) -> int {
    pack_mat_kernel<<<(Kpad * FP + 255) / 256, 256, 0, st>>>(raw->weight[layer], ld, col0, K, F, Kpad, out + off);
    AGX_LAUNCH_CHECK();
    return AGX_OK;
  };
  auto vec = [&](int layer, int n_valid, int n_pad, size_t off) -> int {
    pack_vec_kernel<<<(n_pad + 255) / 256, 256, 0, st>>>(raw->bias[layer], n_valid, n_pad, out + off);
    AGX_LAUNCH_CHECK();
    return AGX_OK;
  };
  int rc = 0;
  rc |= mat(AGX_W_PENC0, d_node, 0, d_node, D_NODE_IN, L.penc0_w); rc |= vec(AGX_W_PENC0, F, FP, L.penc0_b);
  rc |= mat(AGX_W_PENC2, F, 0, F, FP, L.penc2_w);                  rc |= vec(AGX_W_PENC2, F, FP, L.penc2_b);
  rc |= mat(AGX_W_PENC4, F, 0, F, FP, L.penc4_w);                  rc |= vec(AGX_W_PENC4, F, FP, L.penc4_b);
  rc |= mat(AGX_W_RENC0, d_rel, 0, d_rel, D_REL_IN, L.renc0_w);    rc |= vec(AGX_W_RENC0, F, FP, L.renc0_b);
  rc |= mat(AGX_W_RENC2, F, 0, F, FP, L.renc2_w);                  rc |= vec(AGX_W_RENC2, F, FP, L.renc2_b);
  rc |= mat(AGX_W_RENC4, F, 0, F, FP, L.renc4_w);                  rc |= vec(AGX_W_RENC4, F, FP, L.renc4_b);
  rc |= mat(AGX_W_RPROP, 3 * F, 0, F, FP, L.rp_rel_w);             rc |= vec(AGX_W_RPROP, F, FP, L.rp_b);
  rc |= mat(AGX_W_RPROP, 3 * F, F, F, FP, L.rp_recv_w);
  rc |= mat(AGX_W_RPROP, 3 * F, 2 * F, F, FP, L.rp_send_w);
  rc |= mat(AGX_W_PPROP, 2 * F, 0, F, FP, L.pp_enc_w);             rc |= vec(AGX_W_PPROP, F, FP, L.pp_b);
  rc |= mat(AGX_W_PPROP, 2 * F, F, F, FP, L.pp_agg_w);
  rc |= mat(AGX_W_PRED0, F, 0, F, FP, L.pred0_w);                  rc |= vec(AGX_W_PRED0, F, FP, L.pred0_b);
  rc |= mat(AGX_W_PRED1, F, 0, F, FP, L.pred1_w);                  rc |= vec(AGX_W_PRED1, F, FP, L.pred1_b);
  if (rc) return AGX_ERR_CUDA;
  pack_rows_kernel<<<(3 * FP + 255) / 256, 256, 0, st>>>(raw->weight[AGX_W_PRED2], 3, F, out + L.pred2_w);
  AGX_LAUNCH_CHECK();
  if (int rc2 = vec(AGX_W_PRED2, 3, 4, L.pred2_b)) return rc2;
  if (int rc3 = tc_pack(dims, raw, packed, L.total * sizeof(float), st)) return rc3;   // scaled fp16 hi/lo images for the tensor-core path
  return train_pack(dims, raw, packed, st);                                             // plain padded copies for the backward
}

size_t agx_forward_workspace_bytes(const AgxModelDims* dims, int32_t B, int32_t N, int64_t E_cap) {
  (void)dims;
  if (B <= 0 || N <= 0 || E_cap < 0) return 0;
  return agx::fwd_ws_carve(nullptr, (int64_t)B * N, E_cap, nullptr);
}

static int check_graph(const AgxGraphIn* g) {
  AGX_REQUIRE(g, AGX_ERR_ARG, "graph is null");
  AGX_REQUIRE(g->B > 0 && g->N > 0 && g->n_p > 0 && g->n_p <= g->N, AGX_ERR_ARG, "graph: bad sizes B=%d N=%d n_p=%d", g->B, g->N, g->n_p);
  AGX_REQUIRE((int64_t)g->B * g->N < (1ll << 31) - 1, AGX_ERR_ARG, "graph: B*N overflows int32");
  AGX_REQUIRE(g->state && g->attrs && g->action && g->p_instance && g->physics && g->row_ptr, AGX_ERR_ARG, "graph: null tensor pointer");
  AGX_REQUIRE(g->E_cap >= 0 && (g->E_cap == 0 || (g->send && g->recv)), AGX_ERR_ARG, "graph: null edge pointer with E_cap=%lld", (long long)g->E_cap);
  return AGX_OK;
}

int agx_forward(const AgxModelDims* dims, const void* packed_weights, const AgxGraphIn* g, float* pred_pos,
                int64_t pos_stride_b, float* pred_motion, int32_t precision, void* workspace, size_t workspace_bytes,
                agx_stream_t stream) {
  using namespace agx;
  if (int rc = check_dims(dims)) return rc;
  if (int rc = check_graph(g)) return rc;
  AGX_REQUIRE(packed_weights && pred_pos && pred_motion, AGX_ERR_ARG, "forward: null pointer argument");
  AGX_REQUIRE(pos_stride_b >= (int64_t)g->n_p * 3, AGX_ERR_ARG, "forward: pos_stride_b too small");
  return forward_impl(dims, static_cast<const float*>(packed_weights), g, pred_pos, pos_stride_b, pred_motion, precision,
                      workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

size_t agx_rollout_workspace_bytes(const AgxModelDims* dims, int32_t B, int32_t N, int64_t E_cap, int32_t topk) {
  (void)dims;
  if (B <= 0 || N <= 0 || E_cap <= 0 || topk <= 0) return 0;
  const int k = topk < N ? topk : N;
  size_t need = agx::rollout_ws_carve(nullptr, B, N, E_cap, k, nullptr);
  if (B >= 2) {   // the side-by-side plan carves two half-batch workspaces
    const int B0 = B / 2;
    const int64_t E0 = E_cap / B * B0;
    const size_t split = agx::rollout_ws_carve(nullptr, B0, N, E0, k, nullptr) + agx::rollout_ws_carve(nullptr, B - B0, N, E_cap - E0, k, nullptr);
    if (split > need) need = split;
  }
  return need;
}

int agx_rollout(const AgxModelDims* dims, const void* packed_weights, const AgxRolloutIn* r, float* pred_seq,
                int32_t* n_edges_seq, int32_t* status, int32_t precision, void* workspace, size_t workspace_bytes,
                agx_stream_t stream) {
  using namespace agx;
  if (int rc = check_dims(dims)) return rc;
  AGX_REQUIRE(r && packed_weights && pred_seq && status, AGX_ERR_ARG, "rollout: null pointer argument");
  AGX_REQUIRE(r->B > 0 && r->N > 0 && r->n_p > 0 && r->n_p <= r->N && r->n_steps > 0 && r->E_cap > 0 && r->topk > 0, AGX_ERR_ARG,
              "rollout: bad sizes B=%d N=%d n_p=%d T=%d E_cap=%lld", r->B, r->N, r->n_p, r->n_steps, (long long)r->E_cap);
  AGX_REQUIRE(r->state && r->attrs && r->action && r->p_instance && r->physics && r->mask && r->tool_mask && r->thr2, AGX_ERR_ARG,
              "rollout: null tensor pointer");
  AGX_REQUIRE(r->y_mode == AGX_Y_MIN || r->y_mode == AGX_Y_MASKED_MEAN, AGX_ERR_ARG, "rollout: bad y_mode %d", r->y_mode);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int topk = r->topk < r->N ? r->topk : r->N;
  RolloutWs ws;
  rollout_ws_carve(workspace, r->B, r->N, r->E_cap, topk, &ws);
  const size_t need = agx_rollout_workspace_bytes(dims, r->B, r->N, r->E_cap, r->topk);   // covers the side-by-side plan too
  AGX_REQUIRE(workspace && workspace_bytes >= need, AGX_ERR_CAPACITY, "rollout: workspace %zu < %zu bytes", workspace_bytes, need);
  const int T = r->n_steps;
  const int64_t frame = (int64_t)r->N * 3;
  // one model step of the graphs [b0, b0 + Bh) on stream s with workspace w
  auto step = [&](int t, int b0, int Bh, int64_t E_cap_h, const RolloutWs& w, cudaStream_t s, cudaEvent_t after_encoders) -> int {
    float* state = r->state + (size_t)b0 * H_FIX * frame;
    const float* action = r->action + (size_t)b0 * frame;
    const uint8_t* mask = r->mask + (size_t)b0 * r->N;
    // relations on the current positions (forward_dynamics.py:125 for t = 0, :171 afterwards)
    int rc = graph_build_impl(state + (H_FIX - 1) * frame, H_FIX * frame, mask, r->tool_mask + (size_t)b0 * r->N, r->thr2 + b0, Bh, r->N, topk,
                              r->connect_tools_all != 0, AGX_SEM_BATCH, w.row_ptr, w.send, w.recv, E_cap_h,
                              n_edges_seq ? n_edges_seq + (size_t)t * r->B + b0 : w.n_edges, status, w.graph_ws, w.graph_ws_bytes, s);
    if (rc) return rc;
    AgxGraphIn g{Bh, r->N, r->n_p, state, r->attrs + (size_t)b0 * r->N * dims->d_attr, action, r->p_instance + (size_t)b0 * r->n_p,
                 r->physics + (size_t)b0 * dims->d_phys, w.row_ptr, w.send, w.recv, E_cap_h};
    const int64_t stride = (int64_t)T * r->n_p * 3;
    float* pred_t = pred_seq + (size_t)b0 * stride + (size_t)t * r->n_p * 3;
    rc = forward_impl(dims, static_cast<const float*>(packed_weights), &g, pred_t, stride, w.motion, precision, w.fwd_ws, w.fwd_ws_bytes, s,
                      after_encoders, /*reuse_node_products=*/t > 0, /*nfeat_ready=*/t > 0);   // attrs, physics, action: fixed for the rollout
    if (rc) return rc;
    float* nfeat_next = nullptr;                 // tensor-core path: the advance also prepares the next step's history records
    if (precision != AGX_PREC_FP32 && t + 1 < T) {
      FwdWs fws;
      fwd_ws_carve(w.fwd_ws, (int64_t)Bh * r->N, E_cap_h, &fws);
      nfeat_next = fws.nfeat;
    }
    { ProfScope ps(AGX_KIND_ROLLOUT_ADVANCE, s);
      launch_pdl(PDL_SMALL, rollout_advance_kernel, dim3(Bh, (r->N + 255) / 256), 256, 0, s, state, action, mask, pred_t, stride, r->N, r->n_p, r->y_mode, r->gripper_raise, nfeat_next,
                                                g.attrs, g.p_instance); }
    AGX_LAUNCH_CHECK();
    return AGX_OK;
  };

  SplitPlan plan;
  if (rollout_split_plan(r->B, r->N, r->E_cap, topk, precision, &plan)) {
    // Two half-batches side by side (graphs are independent): each on its own stream with the persistent grids sized for half
    // the SMs, the second half started once the first has enqueued its encoder chains, so that one half's tensor-bound phase
    // (node / edge encoders) runs under the other half's HBM-bound phase (aggregate / update) instead of after it.
    SplitStreams* ss = split_streams();
    AGX_REQUIRE(ss, AGX_ERR_CUDA, "rollout: could not create the second stream");
    RolloutWs w0, w1;
    const size_t n0 = rollout_ws_carve(workspace, plan.B0, r->N, plan.E0, topk, &w0);
    rollout_ws_carve(static_cast<char*>(workspace) + n0, r->B - plan.B0, r->N, plan.E1, topk, &w1);
    AGX_CUDA_OK(cudaEventRecord(ss->fork, st));
    AGX_CUDA_OK(cudaStreamWaitEvent(ss->side, ss->fork, 0));
    set_sm_share(plan.sms);
    int rc = AGX_OK;
    for (int t = 0; t < T && !rc; ++t) {
      rc = step(t, 0, plan.B0, plan.E0, w0, st, t == 0 ? ss->offset : nullptr);
      if (!rc && t == 0) rc = cudaStreamWaitEvent(ss->side, ss->offset, 0) == cudaSuccess ? AGX_OK : AGX_ERR_CUDA;
      if (!rc) rc = step(t, plan.B0, r->B - plan.B0, plan.E1, w1, ss->side, nullptr);
    }
    set_sm_share(0);
    if (rc) return rc;
    AGX_CUDA_OK(cudaEventRecord(ss->join, ss->side));
    AGX_CUDA_OK(cudaStreamWaitEvent(st, ss->join, 0));
    return AGX_OK;
  }
  for (int t = 0; t < T; ++t)
    if (int rc = step(t, 0, r->B, r->E_cap, ws, st, nullptr)) return rc;
  return AGX_OK;
}

}  // extern "C"
