// Host interface of the tensor-core model stages (tc_forward.cu), shared by forward.cu (inference / rollout) and train.cu.
#pragma once
#include "common.cuh"

namespace agx {

// Training only: row-major fp32 [rows][FP] copies (rel_in: [E][D_REL_IN], p_in: [rows][D_NODE_IN]) of everything the backward
// reads, written by the same epilogues that feed the next layer.  P_next / agg_f32 / u1 / u2 are per propagation step.
struct TcTrainSave {
  float *p_in, *h1, *h2, *penc;
  float *rel_in, *g1, *g2, *renc;
  float *P_next, *agg_f32, *u1, *u2;
};

struct TcFwdBuffers {
  float* nfeat; float* P; float* A; float* Qr; float* Qs; float* agg; float* C; float* rowmaxP; float* rowmaxA;
  int32_t* agg_exp; float* agg_max;
  float* P0; float* Qr0; float* Qs0; float* rowmaxP0;   // the particle encoder's copies (read-only for the propagation steps)
  float* S0;                                            // A_n + P0
  const TcTrainSave* save = nullptr;                    // non-null: the SAVE instantiations of the chains run (fp32 C only)
  bool agg_f32 = false;                                 // agg holds plain blocked fp32 rows (written by the C16 aggregate, AGX_PREC_TC_MIXED)
};                                                      // instead of the scaled fp16 hi / lo split + agg_exp / agg_max

int tc_edge_aggregate(const AgxGraphIn* g, const TcFwdBuffers& w, bool mixed, bool first, cudaStream_t st);
int tc_nfeat(const AgxGraphIn* g, const TcFwdBuffers& w, cudaStream_t st);
int tc_node_encoder(const AgxGraphIn* g, const float* wts, const PackedLayout& PL, size_t base_bytes, const TcFwdBuffers& w, cudaStream_t st);
int tc_edge_encoder(const AgxGraphIn* g, const float* wts, const PackedLayout& PL, size_t base_bytes, const TcFwdBuffers& w, bool mixed,
                    cudaStream_t st);
int tc_node_update(const AgxGraphIn* g, const float* wts, const PackedLayout& PL, size_t base_bytes, const TcFwdBuffers& w, bool first, bool last,
                   float* pred_pos, int64_t pos_stride_b, float* pred_motion, cudaStream_t st);

}  // namespace agx
