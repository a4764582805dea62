// Host interface of the tensor-core backward chains (tc_backward.inl, compiled as part of tc_forward.cu) used by train.cu.
//
// The backward of DynamicsPredictor.forward (reference: torch autograd over dynamics/gnn/model.py:129-313) as four chain
// kernels on the same tile / slot / split-fp16 machinery as the forward.  Every gradient that a weight-gradient job or a later
// kernel reads leaves the chain as a row-major fp32 row with the ReLU mask already applied.
#pragma once
#include "common.cuh"

namespace agx {

struct TcBwdBuffers {
  // saved forward activations, row-major [rows][FP] / [E][FP]
  const float *u2, *u1, *penc, *h2, *h1, *renc, *g2, *g1;
  // gradients, row-major
  float *dm;                 // [rows][4] gradient with respect to the predicted motion
  float *dU2, *dU1;          // masked gradients of the two hidden head layers
  float *dPre;               // d pre_n (masked) of the propagation step being left: read by the step kernels, written by the head
  float *dPreOut;            // d pre_n of the step being entered (step kernels) / dP_0 (particle encoder kernel)
  float *dA, *dAgg, *dQr, *dQs;
  float *dPenc, *dH2, *dH1;  // masked gradients of the particle encoder outputs
  float *dC, *dE, *dG2, *dG1, *dRel;   // relation side ([E][FP]; dRel [E][D_REL_IN])
  // per-row magnitude bounds (exact power-of-two operand scales are derived from them)
  float *preMax, *aBound, *aggMax, *cBound, *qrMax, *qsMax;
};

// head: d motion -> dU2 -> dU1 -> dP_K ; d pre_{K-1} = dP_K (*) [P_K > 0] ; dA = d pre ; dAgg_{K-1}
int tc_bwd_head(const AgxGraphIn* g, const float* wts, const PackedLayout& PL, size_t base_bytes, const TcBwdBuffers& b, const float* P_K,
                const float* d_pos, const float* d_motion, const float* motion, cudaStream_t st);
// propagation step k >= 1 (after its relation kernels): dP_k = d pre_k + dQr W_recv + dQs W_send ; d pre_{k-1} = dP_k (*) [P_k > 0] ;
// dA += d pre_{k-1} ; dAgg_{k-1}
int tc_bwd_step(const AgxGraphIn* g, const float* wts, size_t base_bytes, const TcBwdBuffers& b, const float* P_k, cudaStream_t st);
// step 0 and the particle encoder: dP_0 ; d penc = dP_0 + dA W_enc ; dH2 ; dH1 (all masked)
int tc_bwd_node_encoder(const AgxGraphIn* g, const float* wts, size_t base_bytes, const TcBwdBuffers& b, cudaStream_t st);
// relation encoder: dC -> dE -> dG2 -> dG1 (masked) -> dRel
int tc_bwd_edge_encoder(const AgxGraphIn* g, const float* wts, size_t base_bytes, const TcBwdBuffers& b, cudaStream_t st);

}  // namespace agx
