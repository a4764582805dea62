/*
 * adaptigraph_b200 — C ABI of the B200-native particle-graph dynamics engine.
 *
 * Drop-in boundary for ONE hot path of Boey-li/AdaptiGraph (reference paths are
 * relative to the reference's src/):
 *
 *   graph construction   dynamics/dataset/graph.py:38-89 (single), :91-156 (batched)
 *   model forward        dynamics/gnn/model.py:129-313
 *   rollout step         planning/forward_dynamics.py:156-197 (and :351-393)
 *   pad / truncate       dynamics/utils.py:37-46, :127-137
 *   particle sampling    dynamics/dataset/graph.py:8-36, dynamics/utils.py:10-24 (the step in front of the path)
 *
 * The reference exposes no FFI on this path (its boundary is three Python call
 * signatures); these entry points are what a ctypes binding on the
 * reference side attaches to.  Conventions:
 *
 *   - extern "C", plain pointers and sizes only.  Every pointer is a DEVICE pointer
 *     unless its name ends in _host.  The caller owns every buffer, including the
 *     workspace; the library holds no persistent device allocations.
 *   - All calls are asynchronous and ordered on `stream` (a cudaStream_t passed as
 *     void*; NULL = the legacy default stream).  No call synchronises the device.
 *   - Return 0 on success; AGX_ERR_* (<0) otherwise, with a thread-local message from
 *     agx_last_error().  Edge-capacity overflow inside a kernel cannot be returned
 *     synchronously: it is reported through the `status` device word (see
 *     agx_graph_build) which the caller reads when it next synchronises.
 *   - Feature rows are fp32 with an internal padded stride AGX_FP (=160 floats,
 *     640 B = five 128-B lines); the padding never leaves the library.
 *   - Relations are CSR by receiver over the flattened node index r = b*N + n:
 *     row_ptr int32 [B*N+1], send int32 [E] (sender id LOCAL to its graph), and the
 *     expanded receiver list recv int32 [E] (flattened id).  Rows appear in the
 *     reference's order: (graph, receiver, sender) ascending (graph.py:151-155).
 */
#ifndef ADAPTIGRAPH_B200_H
#define ADAPTIGRAPH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AGX_VERSION 100          /* 0.1.0 */
#define AGX_FP 160               /* padded feature stride (floats) */
#define AGX_MAX_TOPK 32          /* graph builder keeps one candidate per warp lane */
#define AGX_NFEAT 16             /* per-node relation-input record: hist(12) attr(2) group(1) pad(1) */

#define AGX_OK 0
#define AGX_ERR_ARG (-1)         /* bad argument / unsupported shape */
#define AGX_ERR_CAPACITY (-2)    /* workspace or edge capacity too small (pad_torch's failure, utils.py:37-46) */
#define AGX_ERR_CUDA (-3)        /* CUDA runtime error, see agx_last_error() */

/* graph-builder semantics */
#define AGX_SEM_BATCH 0          /* construct_edges_from_states_batch, graph.py:91-156 */
#define AGX_SEM_SINGLE 1         /* construct_edges_from_states,       graph.py:38-89  */

/* rollout tool-height rule */
#define AGX_Y_MIN 0              /* y = min_n pred[n].y           forward_dynamics.py:163 */
#define AGX_Y_MASKED_MEAN 1      /* y = masked mean of pred[n].y  forward_dynamics.py:359 */

/* arithmetic of the dense 150x150 layers */
#define AGX_PREC_FP32 0          /* exact fp32 FFMA tiles */
#define AGX_PREC_TC_F16X3 1      /* tcgen05 kind::f16 on per-row power-of-two scaled fp16 hi/lo splits of both operands
                                    (3 MMAs per K step, 22 significant bits), fp32 accumulate in tensor memory */
#define AGX_PREC_TC_MIXED 2      /* same, with the precision budget spent where the 1e-4 rollout tolerance allows: the relation chain
                                    (relation_encoder.model.2/.4, relation part of the propagator) rounds its activations to fp16
                                    (2 MMAs per K step) and the per-relation term C is kept as 16-bit block fixed point; every
                                    particle-side layer stays at 3 MMAs.  Rollout RMSE ~3.5e-6 instead of ~6e-7. */

#if defined(__GNUC__)
#define AGX_API __attribute__((visibility("default")))
#else
#define AGX_API
#endif

typedef void* agx_stream_t;

/* Model dimensions (DynamicsPredictor.__init__, model.py:77-122). */
typedef struct AgxModelDims {
  int32_t F;        /* nf_particle = nf_relation = nf_effect (<= 160; 150 in every shipped config) */
  int32_t n_his;    /* history frames H (4) */
  int32_t d_attr;   /* attr_dim (2) */
  int32_t d_phys;   /* number of physics params in use (1) */
  int32_t d_act;    /* action_dim (3) */
  int32_t pstep;    /* propagation steps K */
} AgxModelDims;

/* Reference-layout parameters, row-major [out][in] fp32 exactly as in the state_dict
 * (model.py:103-122).  Index order: */
enum {
  AGX_W_PENC0 = 0, AGX_W_PENC2, AGX_W_PENC4,   /* particle_encoder.model.{0,2,4} */
  AGX_W_RENC0, AGX_W_RENC2, AGX_W_RENC4,       /* relation_encoder.model.{0,2,4} */
  AGX_W_PPROP,                                 /* particle_propagator.linear (F x 2F) */
  AGX_W_RPROP,                                 /* relation_propagator.linear (F x 3F) */
  AGX_W_PRED0, AGX_W_PRED1, AGX_W_PRED2,       /* non_rigid_predictor.linear_{0,1,2} */
  AGX_NUM_LAYERS
};
typedef struct AgxWeights {
  const float* weight[AGX_NUM_LAYERS];
  const float* bias[AGX_NUM_LAYERS];
} AgxWeights;

/* One batch of graphs (the reference's graph dict, forward_dynamics.py:130-147). */
typedef struct AgxGraphIn {
  int32_t B, N, n_p;       /* graphs, particles per graph (object + tool), object particles */
  const float* state;      /* (B, H, N, 3) */
  const float* attrs;      /* (B, N, d_attr) */
  const float* action;     /* (B, N, 3) */
  const float* p_instance; /* (B, n_p) — single instance column (max_n = 1 in every shipped config) */
  const float* physics;    /* (B, d_phys) */
  const int32_t* row_ptr;  /* (B*N + 1) */
  const int32_t* send;     /* (E) */
  const int32_t* recv;     /* (E) */
  int64_t E_cap;           /* upper bound on E used to size the workspace (E itself is row_ptr[B*N], read on device) */
} AgxGraphIn;

AGX_API int agx_version(void);
AGX_API const char* agx_last_error(void);

/* ---- weights: pad / transpose / split the propagator matrices once per parameter update */
AGX_API size_t agx_packed_weights_bytes(const AgxModelDims* dims);
AGX_API int agx_pack_weights(const AgxModelDims* dims, const AgxWeights* raw, void* packed, agx_stream_t stream);

/* ---- graph construction (replaces graph.py:38-156) */
AGX_API size_t agx_graph_workspace_bytes(int32_t B, int32_t N, int32_t topk);
/* pos (B,N,3); mask, tool_mask (B,N) uint8; thr2 (B) = squared radius per graph.
 * Writes row_ptr (B*N+1), send/recv (up to cap entries), n_edges (B) per-graph counts.
 * status: one int32 device word, OR-ed with 1 if the total edge count exceeded cap
 * (entries beyond cap are dropped, row_ptr still holds the true counts). */
AGX_API int agx_graph_build(const float* pos, const uint8_t* mask, const uint8_t* tool_mask, const float* thr2,
                    int32_t B, int32_t N, int32_t topk, int32_t connect_tools_all, int32_t semantics,
                    int32_t* row_ptr, int32_t* send, int32_t* recv, int64_t cap, int32_t* n_edges,
                    int32_t* status, void* workspace, size_t workspace_bytes, agx_stream_t stream);

/* Dense one-hot relations (B, n_rel, N) -> per-row receiver / sender ids (-1 on all-zero rows),
 * the conversion the drop-in forward() applies when it is handed Rr / Rs (model.py:129). */
AGX_API int agx_onehot_to_ids(const float* R, int32_t B, int32_t n_rel, int32_t N, int32_t* ids, agx_stream_t stream);
/* CSR -> dense one-hot rows (graph.py:152-155): Rr, Rs (B, n_rel, N) must be zero-filled by the caller. */
AGX_API int agx_edges_to_onehot(const int32_t* row_ptr, const int32_t* send, int32_t B, int32_t N, int32_t n_rel,
                        float* Rr, float* Rs, agx_stream_t stream);

/* ---- farthest-point sampling: the particle sampler in front of the path.
 * radius < 0  : dgl.geometry.farthest_point_sampler(pos, max_samples, start_idx) as called at graph.py:11-12 and
 *               perception.py:271 — exactly max_samples picks per cloud (squared fp32 distances, first index on ties).
 * radius >= 0 : fps_rad_idx(pcd, radius), utils.py:10-24 — picks until every point is within `radius` of a pick
 *               (fp32 norms, compared with the double `radius`), at most max_samples.
 * pos (B,N,3); n_points (B) valid prefix length of every cloud or NULL (= N); start_idx (B) first pick;
 * idx_out (B, max_samples) picks in selection order; n_out (B) number of picks.  Clouds of up to 12800 points run in one CTA's
 * shared memory; larger ones (N <= 204800) on a thread-block cluster of 2..16 CTAs exchanging their arg-max through distributed
 * shared memory, with identical picks. */
AGX_API int agx_fps(const float* pos, const int32_t* n_points, int32_t B, int32_t N, int32_t max_samples,
            const int32_t* start_idx, double radius, int32_t* idx_out, int32_t* n_out, agx_stream_t stream);
/* The radius form with one radius per cloud (device array of B doubles, each >= 0): a training batch draws
 * fps_radius ~ U(fps_radius_range) per sample (graph.py:17-20), and the batched loader thins all its samples in one launch. */
AGX_API int agx_fps_radii(const float* pos, const int32_t* n_points, int32_t B, int32_t N, int32_t max_samples,
                  const int32_t* start_idx, const double* radii, int32_t* idx_out, int32_t* n_out, agx_stream_t stream);

/* ---- chamfer distance of the MPC error term (planning/losses.py:4-10, used at plan.py:36 / :146 on the rollout's frames):
 * out[b] = mean_m min_n |x[b,n] - y[m]| + mean_n min_m |x[b,n] - y[m]|.  x (B,N,3); y (M,3) shared by every sample
 * (y_batched = 0, the planner's target) or (B,M,3) (y_batched = 1); out (B).  N + M <= 17066. */
AGX_API int agx_chamfer(const float* x, const float* y, int32_t B, int32_t N, int32_t M, int32_t y_batched, float* out,
                agx_stream_t stream);

/* ---- the planner's reward tail in one launch: running_cost (planning/plan.py:27-59) with its error term (chamfer to target points,
 * losses.py:4-10, or box_loss to a target box, :25-35; plan.py:146 / :155), its collision penalty (rope / cloth / granular,
 * losses.py:37-92; plan.py:160-165) and the workspace-box penalty (plan.py:41-51).  state (bsz, L, n, 3) predicted particles per
 * look-ahead step, action (bsz, L, action_dim >= 3), state_cur (n, 3), bbox (2, 2), target (M, 3) points or (2, 2) box,
 * reward (bsz) = -2 / (max error + 1e-6) * error[:, -1] - 5 * mean_l penalty - 5 * mean_l box penalty.  n + M <= 17066. */
#define AGX_ERROR_CHAMFER 0
#define AGX_ERROR_BOX 1
#define AGX_PENALTY_ROPE 0
#define AGX_PENALTY_CLOTH 1
#define AGX_PENALTY_GRANULAR 2
AGX_API size_t agx_running_cost_workspace_bytes(int32_t bsz, int32_t L);
/* workspace: agx_running_cost_workspace_bytes(bsz, L) bytes, ZERO-filled before the first call (the kernel leaves it ready for the next) */
AGX_API int agx_running_cost(const float* state, const float* action, int32_t action_dim, const float* state_cur, const float* bbox,
                     int32_t error_mode, const float* target, int32_t M, int32_t penalty_mode, float sim_real_ratio, int32_t bsz,
                     int32_t L, int32_t n, void* workspace, size_t workspace_bytes, float* reward, agx_stream_t stream);

/* ---- model forward (replaces model.py:129-313) */
AGX_API size_t agx_forward_workspace_bytes(const AgxModelDims* dims, int32_t B, int32_t N, int64_t E_cap);
/* pred_pos, pred_motion: (B, n_p, 3).  pos_stride_b: floats between consecutive graphs in
 * pred_pos (n_p*3 for a dense (B,n_p,3) tensor; T*n_p*3 when writing step t of a (B,T,n_p,3) rollout). */
AGX_API int agx_forward(const AgxModelDims* dims, const void* packed_weights, const AgxGraphIn* g,
                float* pred_pos, int64_t pos_stride_b, float* pred_motion,
                int32_t precision, void* workspace, size_t workspace_bytes, agx_stream_t stream);

/* ---- autoregressive rollout (replaces the loop of forward_dynamics.py:156-197) */
typedef struct AgxRolloutIn {
  int32_t B, N, n_p;
  float* state;              /* (B, H, N, 3): in = initial history, out = history after the last step */
  const float* attrs;        /* (B, N, d_attr) */
  const float* action;       /* (B, N, 3) — constant over the rollout (forward_dynamics.py:180) */
  const float* p_instance;   /* (B, n_p) */
  const float* physics;      /* (B, d_phys) */
  const uint8_t* mask;       /* (B, N) state_mask */
  const uint8_t* tool_mask;  /* (B, N) eef_mask */
  const float* thr2;         /* (B) */
  int32_t topk, connect_tools_all;
  int32_t n_steps;           /* T */
  int32_t y_mode;            /* AGX_Y_MIN / AGX_Y_MASKED_MEAN */
  float gripper_raise;       /* 0.01*sim_real_ratio when gripper_enable (forward_dynamics.py:167-168), else 0 */
  int64_t E_cap;             /* max relations per batch the workspace is sized for (B * max_nR) */
} AgxRolloutIn;
AGX_API size_t agx_rollout_workspace_bytes(const AgxModelDims* dims, int32_t B, int32_t N, int64_t E_cap, int32_t topk);
/* pred_seq (B, T, n_p, 3); n_edges_seq (T, B) int32 per-step per-graph relation counts (nullable);
 * status as in agx_graph_build. */
AGX_API int agx_rollout(const AgxModelDims* dims, const void* packed_weights, const AgxRolloutIn* r,
                float* pred_seq, int32_t* n_edges_seq, int32_t* status,
                int32_t precision, void* workspace, size_t workspace_bytes, agx_stream_t stream);

/* ---- training (replaces torch autograd through model.py:129-313 in dynamics/train/train.py:90-112)
 * agx_forward_train is agx_forward in exact fp32 that additionally keeps every activation the backward needs in the
 * caller's `saved` buffer.  agx_backward consumes it: parameter gradients are ACCUMULATED (+=) into reference-layout
 * tensors (null entries are skipped), d_state (B,H,N,3) is accumulated too (nullable).  send_ptr (B*N+1) / send_perm (E)
 * list the same relations grouped by (flattened) sender, in a fixed order, so the sender-side reductions are
 * deterministic.  pred_motion is the forward's output (needed for the clamp mask, model.py:309). */
typedef struct AgxWeightGrads {
  float* weight[AGX_NUM_LAYERS];
  float* bias[AGX_NUM_LAYERS];
} AgxWeightGrads;
AGX_API size_t agx_train_saved_bytes(const AgxModelDims* dims, int32_t B, int32_t N, int64_t E_cap);
AGX_API size_t agx_train_scratch_bytes(const AgxModelDims* dims, int32_t B, int32_t N, int64_t E_cap);
/* Where the forward activations live inside `saved` (for parity tests that compare ReLU activity patterns with the reference's;
 * the backward itself needs nothing from the caller).  Fills out[0 .. n-1] with BYTE offsets, in this order:
 *   0 h1, 1 h2, 2 penc (= particle effect 0), 3 g1, 4 g2, 5 renc, 6 C, 7 u1, 8 u2, then per propagation step k = 0 .. pstep-1:
 *   9+4k particle effect k+1, 10+4k agg_k, 11+4k Qr_k, 12+4k Qs_k.
 * Rows are fp32 with a stride of AGX_FP floats (particles: B*N rows, relations: E_cap rows), except C / Qr / Qs when the return
 * value is 1: those then use the blocked layout of the tensor-core path -- [row / 128][piece][row % 128][w] floats with nine
 * 16-column pieces and a tenth of 8 columns (columns 0..151; 152..159 are padding and not stored).  Returns 0 (all row-major),
 * 1 (blocked C / Qr / Qs) or a negative error code. */
AGX_API int agx_train_saved_offsets(const AgxModelDims* dims, int32_t B, int32_t N, int64_t E_cap, int64_t* out, int32_t n);
AGX_API int agx_forward_train(const AgxModelDims* dims, const void* packed_weights, const AgxGraphIn* g, float* pred_pos,
                              int64_t pos_stride_b, float* pred_motion, void* saved, size_t saved_bytes, agx_stream_t stream);
AGX_API int agx_backward(const AgxModelDims* dims, const void* packed_weights, const AgxGraphIn* g, const void* saved,
                         const int32_t* send_ptr, const int32_t* send_perm, const float* pred_motion, const float* d_pred_pos,
                         const float* d_pred_motion, const AgxWeightGrads* grads, float* d_state, void* scratch,
                         size_t scratch_bytes, agx_stream_t stream);

/* ---- optimiser step (replaces torch.optim.Adam of dynamics/train/train.py:63, :115): Adam over one flat
 * fp32 bucket of n parameters (params, grads, exp_avg, exp_avg_sq: 16-byte aligned device arrays of n floats).  Gradients are
 * multiplied by grad_scale first (1 / world size after a summing all-reduce).  `step` is a device int32 holding the number of
 * steps taken so far; it is read for the bias corrections and incremented on the stream, so a captured CUDA graph can be replayed.
 * The hyper-parameters are doubles because torch derives its scalars (1 - beta, lr / (1 - beta1^t), sqrt(1 - beta2^t)) in double. */
AGX_API int agx_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1,
                  double beta2, double eps, float grad_scale, int32_t* step, agx_stream_t stream);

/* ---- per-kernel timing for bench.py's roofline block.  When enabled, every kernel launched by the
 * calling thread is bracketed by CUDA events on its launch stream; agx_profile_read synchronises on
 * those events and ADDS elapsed milliseconds / launch counts per kernel kind into ms[] / count[]
 * (arrays of AGX_NUM_KINDS), then forgets the recorded events. */
enum {
  AGX_KIND_GRAPH_TOOLS = 0, AGX_KIND_GRAPH_KNN, AGX_KIND_GRAPH_SCAN, AGX_KIND_GRAPH_FILL,
  AGX_KIND_NODE_ENCODER, AGX_KIND_EDGE_ENCODER, AGX_KIND_EDGE_AGGREGATE, AGX_KIND_NODE_UPDATE,
  AGX_KIND_NODE_HEAD, AGX_KIND_ROLLOUT_ADVANCE, AGX_KIND_OTHER, AGX_KIND_GRAPH_SORT, AGX_NUM_KINDS
};
AGX_API int agx_profile_enable(int32_t on);
AGX_API int agx_profile_read(double* ms, int64_t* count);
AGX_API const char* agx_kind_name(int32_t kind);

/* Cumulative number of kernel launches issued through this library by the calling thread
 * (bench.py differences it around the timed region for gpu_launches). */
AGX_API int64_t agx_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* ADAPTIGRAPH_B200_H */
