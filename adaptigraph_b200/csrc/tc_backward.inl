// Backward chains of the training path (see tc_backward.cuh).  Included at the end of tc_forward.cu: same tile / slot machinery,
// weight loader and MMA roles as the forward chains (tc_chain.cuh), with the TRANSPOSED weight images TT_* as B operands
// (dX = dY * W).  Differences from the forward epilogues:
//   * no ReLU and no bias column: a layer's result is multiplied by the 0/1 mask of the forward activation it passes through
//     (read from the saved fp32 row: act > 0), stored once as a row-major fp32 row -- the operand of that layer's weight-gradient
//     job and of nothing else -- and re-split as the next layer's A;
//   * operand scales come from per-row bounds that the producing kernels leave next to the rows (exact row maxima where a row
//     is stored anyway, sums of such maxima for the accumulated dA / dC), propagated through the layers with the row-sum norm of
//     the transposed matrices;
//   * where a gradient is a sum of two products (dP_k = d pre + dQr W_recv + dQs W_send) the first product parks in the row's
//     own buffer and the epilogue swaps the second operand into A (epi_layer_to_a's side hook may replace what is split).
namespace agx {
namespace tc {

struct BwdArgs {
  int B, N, n_p;
  int64_t rows, E_cap;
  const int32_t* row_ptr; const int32_t* recv;
  const uint8_t* blob; TcLayout L;
  TcBwdBuffers b;
  const float* P_act;     // forward particle effects whose ReLU this kernel crosses (P_K for the head, P_k for step k)
  const float* head_w;    // fp32 [3][FP] + 4 bias floats (head only)
  const float* d_pos; const float* d_motion; const float* motion;
};

__device__ __forceinline__ void zero16(float (&v)[HW]) {
#pragma unroll
  for (int i = 0; i < HW; ++i) v[i] = 0.f;
}
// v *= [act[row][col0 .. col0+15] > 0]
__device__ __forceinline__ void mask16(const float* act, int64_t row, int col0, float (&v)[HW]) {
  const float4* p = reinterpret_cast<const float4*>(act + row * FP + col0);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 m = __ldg(p + q);
    v[4 * q] = m.x > 0.f ? v[4 * q] : 0.f;
    v[4 * q + 1] = m.y > 0.f ? v[4 * q + 1] : 0.f;
    v[4 * q + 2] = m.z > 0.f ? v[4 * q + 2] : 0.f;
    v[4 * q + 3] = m.w > 0.f ? v[4 * q + 3] : 0.f;
  }
}
__device__ __forceinline__ void row_load16(const float* base, int64_t row, int col0, float (&v)[HW]) {
  const float* p = base + row * FP + col0;
  float t[8];
  ldg256(p, t);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = t[i];
  ldg256(p + 8, t);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[8 + i] = t[i];
}
__device__ __forceinline__ void row_add16(const float* base, int64_t row, int col0, float (&v)[HW]) {   // v += stored row piece
  const float* p = base + row * FP + col0;
  float t[8];
  ldg256(p, t);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] += t[i];
  ldg256(p + 8, t);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[8 + i] += t[i];
}
// *row piece += v (v is left untouched)
__device__ __forceinline__ void row_accumulate16(float* base, int64_t row, int col0, const float (&v)[HW]) {
  float t[HW];
#pragma unroll
  for (int i = 0; i < HW; ++i) t[i] = v[i];
  row_add16(base, row, col0, t);
  row_store16(base, row, col0, t);
}

// input producer: row `row` of a row-major fp32 matrix -> the slot's A, scaled from its bound
__device__ __forceinline__ void bwd_produce_rows(const Shared& sh, EpiCtx& cx, const float* src, int64_t row, bool valid, float bound) {
  cx.e_in = scale_exp(bound);
  cx.bound_in = bound;
  const float sc = exp2i(cx.e_in);
#pragma unroll
  for (int c = 0; c < NCHUNK; ++c) {
    float v[HW];
    if (valid) row_load16(src, row, 32 * c + HW * cx.half, v);
    else zero16(v);
    epi_store_a(cx, c, v, sc);
  }
  epi_signal(cx, &sh.bar_in[cx.slot]);
}

// hidden backward layer: next A = (acc * unscale [+ add row]) (*) [act > 0], the masked row also stored to `out`.
// Returns the row's actual maximum (both column halves).
__device__ __forceinline__ float bwd_hidden(const Shared& sh, EpiCtx& cx, const float4 m, float extra_bound, const float* add, const float* act,
                                            float* out, int64_t row, bool valid) {
  const float bound_next = cx.bound_in * m.y + extra_bound;
  const int e_next = scale_exp(bound_next);
  float mx = epi_layer_to_a<true, false>(
      sh, cx, exp2i(-cx.e_in) * m.x, exp2i(e_next),
      [&](int, int col0, float (&v)[HW]) {
        if (valid) {
          if (add) row_add16(add, row, col0, v);
          mask16(act, row, col0, v);
        } else {
          zero16(v);
        }
      },
      [&](int, int col0, float (&v)[HW]) {
        if (valid) row_store16(out, row, col0, v);
      });
  mx = epi_exchange<true>(sh, cx, mx);
  cx.bound_in = mx;
  cx.e_in = e_next;
  return mx;
}

// last layer of a chain: fp32 rows out (optionally masked), returns the row maximum
__device__ __forceinline__ float bwd_out(const Shared& sh, EpiCtx& cx, const float4 m, const float* act, float* out, int64_t row, bool valid) {
  float mx = 0.f;
  epi_layer_out<false>(sh, cx, exp2i(-cx.e_in) * m.x, [&](int, int col0, float (&v)[HW]) {
    if (valid) {
      if (act) mask16(act, row, col0, v);
      mx = max16(v, mx);
      row_store16(out, row, col0, v);
    }
  });
  return epi_exchange<true>(sh, cx, mx);
}

// dP = (d pre + dQr W_recv) + dQs W_send, first half: d pre + layer TT_RP_RECV's result parks in dPreOut, dQs becomes the next A.
// Returns the bound on the parked rows.
__device__ __forceinline__ float bwd_recv_then_send(const Shared& sh, EpiCtx& cx, const float4 m_recv, const BwdArgs& a, int64_t r, bool valid) {
  const float bound_r = cx.bound_in;
  const float bound_s = valid ? a.b.qsMax[r] : 0.f;
  const int e_s = scale_exp(bound_s);
  epi_layer_to_a<true, false>(
      sh, cx, exp2i(-cx.e_in) * m_recv.x, exp2i(e_s), NoExtra{},
      [&](int, int col0, float (&v)[HW]) {
        if (valid) {
          row_add16(a.b.dPre, r, col0, v);
          row_store16(a.b.dPreOut, r, col0, v);
          row_load16(a.b.dQs, r, col0, v);
        } else {
          zero16(v);
        }
      });
  cx.e_in = e_s;
  cx.bound_in = bound_s;
  return (valid ? a.b.preMax[r] : 0.f) + bound_r * m_recv.y;
}

// ------------------------------------------------------------------------------------ head
__global__ void __launch_bounds__(THREADS, 1) tc_bwd_head_kernel(const BwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  constexpr LayerStep prog[3] = {{TT_PRED1, 10, IN_PRODUCER, 3}, {TT_PRED0, 10, IN_EPILOGUE, 3}, {TT_PP_AGG, 10, IN_EPILOGUE, 3}};
  Shared sh;
  float4 meta[3];
  const uint32_t tmem_base = chain_setup(sh, smem_raw, prog, a.blob, a.L, meta);
  for (int i = threadIdx.x; i < 3 * FP + 4; i += THREADS) sh.head_w[i] = a.head_w[i];
  __syncthreads();
  float max_w = 0.f;   // max |V2|: bounds the rows of d motion * V2
  for (int i = threadIdx.x; i < 3 * FP; i += THREADS) max_w = fmaxf(max_w, fabsf(sh.head_w[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) max_w = fmaxf(max_w, __shfl_xor_sync(0xffffffffu, max_w, o));
  if ((threadIdx.x & 31) == 0) sh.xchg[threadIdx.x >> 5] = max_w;
  __syncthreads();
  for (int i = 0; i < THREADS / 32; ++i) max_w = fmaxf(max_w, sh.xchg[i]);
  __syncthreads();
  const int n_tiles = (int)((a.rows + TILE - 1) / TILE);
  const int warp = threadIdx.x >> 5;
  if (warp >= EPI_WARPS) {
    reg_dealloc_other();
    if (warp == LOAD_WARP) loader_role(sh, prog, a.blob, a.L, n_tiles, 21);
    else if (warp < LOAD_WARP) mma_role(sh, prog, warp - MMA_WARP0, tmem_base, n_tiles, 21);
  } else {
    reg_alloc_epilogue();
    EpiCtx cx = make_ctx(tmem_base);
    int tile = slot_tile(0, cx.slot, n_tiles);
    for (int k = 0; tile >= 0; ++k) {
      const int64_t r = (int64_t)tile * TILE + cx.row;
      const bool valid = r < a.rows;
      // ---- producer: d motion (clamp of pred_pos, model.py:309) -> dU2 = (d motion * V2) (*) [u2 > 0]
      float dm0 = 0.f, dm1 = 0.f, dm2 = 0.f;
      if (valid) {
        const int b = (int)(r / a.N), n = (int)(r - (int64_t)b * a.N);
        if (n < a.n_p) {
          const size_t o3 = ((size_t)b * a.n_p + n) * 3;
          const float mo0 = a.motion[o3], mo1 = a.motion[o3 + 1], mo2 = a.motion[o3 + 2];
          dm0 = (a.d_motion ? a.d_motion[o3] : 0.f) + ((a.d_pos && fabsf(mo0) <= MOTION_CLAMP) ? a.d_pos[o3] : 0.f);
          dm1 = (a.d_motion ? a.d_motion[o3 + 1] : 0.f) + ((a.d_pos && fabsf(mo1) <= MOTION_CLAMP) ? a.d_pos[o3 + 1] : 0.f);
          dm2 = (a.d_motion ? a.d_motion[o3 + 2] : 0.f) + ((a.d_pos && fabsf(mo2) <= MOTION_CLAMP) ? a.d_pos[o3 + 2] : 0.f);
        }
        if (cx.half == 0) *reinterpret_cast<float4*>(a.b.dm + r * 4) = make_float4(dm0, dm1, dm2, 0.f);
      }
      {
        const float bound = (fabsf(dm0) + fabsf(dm1) + fabsf(dm2)) * max_w;
        cx.e_in = scale_exp(bound);
        cx.bound_in = bound;
        const float sc = exp2i(cx.e_in);
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c) {
          const int col0 = 32 * c + HW * cx.half;
          float v[HW];
#pragma unroll
          for (int i = 0; i < HW; ++i)
            v[i] = dm0 * lds32(sh.head_w + col0 + i) + dm1 * lds32(sh.head_w + FP + col0 + i) + dm2 * lds32(sh.head_w + 2 * FP + col0 + i);
          if (valid) {
            mask16(a.b.u2, r, col0, v);
            row_store16(a.b.dU2, r, col0, v);
          } else {
            zero16(v);
          }
          epi_store_a(cx, c, v, sc);
        }
        epi_signal(cx, &sh.bar_in[cx.slot]);
      }
      bwd_hidden(sh, cx, meta[0], 0.f, nullptr, a.b.u1, a.b.dU1, r, valid);                    // dU1
      const float pm = bwd_hidden(sh, cx, meta[1], 0.f, nullptr, a.P_act, a.b.dPre, r, valid);  // d pre_{K-1} = dP_K (*) [P_K > 0]
      if (valid) {   // dA starts as this step's d pre (the row this thread just stored: L1 / L2 hit)
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c) {
          const int col0 = 32 * c + HW * cx.half;
          float v[HW];
          row_load16(a.b.dPre, r, col0, v);
          row_store16(a.b.dA, r, col0, v);
        }
        if (cx.half == 0) { a.b.preMax[r] = pm; a.b.aBound[r] = pm; }
      }
      const float am = bwd_out(sh, cx, meta[2], nullptr, a.b.dAgg, r, valid);                  // dAgg_{K-1}
      if (valid && cx.half == 0) { a.b.aggMax[r] = am; a.b.cBound[r] = am; }
      tile = slot_tile(k + 1, cx.slot, n_tiles);
    }
  }
  chain_teardown(tmem_base);
}

// ------------------------------------------------------------------------------------ propagation step k >= 1
__global__ void __launch_bounds__(THREADS, 1) tc_bwd_step_kernel(const BwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  constexpr LayerStep prog[3] = {{TT_RP_RECV, 10, IN_PRODUCER, 3}, {TT_RP_SEND, 10, IN_EPILOGUE, 3}, {TT_PP_AGG, 10, IN_EPILOGUE, 3}};
  Shared sh;
  float4 meta[3];
  const uint32_t tmem_base = chain_setup(sh, smem_raw, prog, a.blob, a.L, meta);
  const int n_tiles = (int)((a.rows + TILE - 1) / TILE);
  const int warp = threadIdx.x >> 5;
  if (warp >= EPI_WARPS) {
    reg_dealloc_other();
    if (warp == LOAD_WARP) loader_role(sh, prog, a.blob, a.L, n_tiles, 22);
    else if (warp < LOAD_WARP) mma_role(sh, prog, warp - MMA_WARP0, tmem_base, n_tiles, 22);
  } else {
    reg_alloc_epilogue();
    EpiCtx cx = make_ctx(tmem_base);
    int tile = slot_tile(0, cx.slot, n_tiles);
    for (int k = 0; tile >= 0; ++k) {
      const int64_t r = (int64_t)tile * TILE + cx.row;
      const bool valid = r < a.rows;
      bwd_produce_rows(sh, cx, a.b.dQr, r, valid, valid ? a.b.qrMax[r] : 0.f);
      const float bound_t = bwd_recv_then_send(sh, cx, meta[0], a, r, valid);
      // dP_k = parked + dQs W_send ; d pre_{k-1} = dP_k (*) [P_k > 0] -> dPreOut (in place of the parked rows), dA += d pre_{k-1}
      const float pm = bwd_hidden(sh, cx, meta[1], bound_t, a.b.dPreOut, a.P_act, a.b.dPreOut, r, valid);
      if (valid) {
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c) {
          const int col0 = 32 * c + HW * cx.half;
          float v[HW];
          row_load16(a.b.dPreOut, r, col0, v);
          row_accumulate16(a.b.dA, r, col0, v);
        }
        if (cx.half == 0) { a.b.preMax[r] = pm; a.b.aBound[r] += pm; }
      }
      const float am = bwd_out(sh, cx, meta[2], nullptr, a.b.dAgg, r, valid);   // dAgg_{k-1}
      if (valid && cx.half == 0) { a.b.aggMax[r] = am; a.b.cBound[r] += am; }
      tile = slot_tile(k + 1, cx.slot, n_tiles);
    }
  }
  chain_teardown(tmem_base);
}

// ------------------------------------------------------------------------------------ step 0 + particle encoder
__global__ void __launch_bounds__(THREADS, 1) tc_bwd_node_encoder_kernel(const BwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  constexpr LayerStep prog[5] = {{TT_RP_RECV, 10, IN_PRODUCER, 3}, {TT_RP_SEND, 10, IN_EPILOGUE, 3}, {TT_PP_ENC, 10, IN_EPILOGUE, 3},
                                 {TT_PENC4, 10, IN_EPILOGUE, 3}, {TT_PENC2, 10, IN_EPILOGUE, 3}};
  Shared sh;
  float4 meta[5];
  const uint32_t tmem_base = chain_setup(sh, smem_raw, prog, a.blob, a.L, meta);
  const int n_tiles = (int)((a.rows + TILE - 1) / TILE);
  const int warp = threadIdx.x >> 5;
  if (warp >= EPI_WARPS) {
    reg_dealloc_other();
    if (warp == LOAD_WARP) loader_role(sh, prog, a.blob, a.L, n_tiles, 23);
    else if (warp < LOAD_WARP) mma_role(sh, prog, warp - MMA_WARP0, tmem_base, n_tiles, 23);
  } else {
    reg_alloc_epilogue();
    EpiCtx cx = make_ctx(tmem_base);
    int tile = slot_tile(0, cx.slot, n_tiles);
    for (int k = 0; tile >= 0; ++k) {
      const int64_t r = (int64_t)tile * TILE + cx.row;
      const bool valid = r < a.rows;
      bwd_produce_rows(sh, cx, a.b.dQr, r, valid, valid ? a.b.qrMax[r] : 0.f);
      const float bound_t = bwd_recv_then_send(sh, cx, meta[0], a, r, valid);
      // dP_0 = parked + dQs W_send -> dPreOut ; the accumulated dA becomes the next A
      float bound_p0;
      {
        const float4 m = meta[1];
        bound_p0 = bound_t + cx.bound_in * m.y;
        const float bound_a = valid ? a.b.aBound[r] : 0.f;
        const int e_a = scale_exp(bound_a);
        epi_layer_to_a<true, false>(
            sh, cx, exp2i(-cx.e_in) * m.x, exp2i(e_a), NoExtra{},
            [&](int, int col0, float (&v)[HW]) {
              if (valid) {
                row_add16(a.b.dPreOut, r, col0, v);
                row_store16(a.b.dPreOut, r, col0, v);
                row_load16(a.b.dA, r, col0, v);
              } else {
                zero16(v);
              }
            });
        cx.e_in = e_a;
        cx.bound_in = bound_a;
      }
      bwd_hidden(sh, cx, meta[2], bound_p0, a.b.dPreOut, a.b.penc, a.b.dPenc, r, valid);   // d penc = (dP_0 + dA W_enc) (*) [penc > 0]
      bwd_hidden(sh, cx, meta[3], 0.f, nullptr, a.b.h2, a.b.dH2, r, valid);
      bwd_out(sh, cx, meta[4], a.b.h1, a.b.dH1, r, valid);
      tile = slot_tile(k + 1, cx.slot, n_tiles);
    }
  }
  chain_teardown(tmem_base);
}

// ------------------------------------------------------------------------------------ relation encoder
__global__ void __launch_bounds__(THREADS, 1) tc_bwd_edge_encoder_kernel(const BwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  constexpr LayerStep prog[4] = {{TT_RP_REL, 10, IN_PRODUCER, 3}, {TT_RENC4, 10, IN_EPILOGUE, 3}, {TT_RENC2, 10, IN_EPILOGUE, 3},
                                 {TT_RENC0, 10, IN_EPILOGUE, 3}};
  Shared sh;
  float4 meta[4];
  const uint32_t tmem_base = chain_setup(sh, smem_raw, prog, a.blob, a.L, meta);
  const int64_t E = min((int64_t)a.row_ptr[a.rows], a.E_cap);
  const int n_tiles = (int)((E + TILE - 1) / TILE);
  const int warp = threadIdx.x >> 5;
  if (warp >= EPI_WARPS) {
    reg_dealloc_other();
    if (warp == LOAD_WARP) loader_role(sh, prog, a.blob, a.L, n_tiles, 24);
    else if (warp < LOAD_WARP) mma_role(sh, prog, warp - MMA_WARP0, tmem_base, n_tiles, 24);
  } else {
    reg_alloc_epilogue();
    EpiCtx cx = make_ctx(tmem_base);
    int tile = slot_tile(0, cx.slot, n_tiles);
    for (int k = 0; tile >= 0; ++k) {
      const int64_t e = (int64_t)tile * TILE + cx.row;
      const bool valid = e < E;
      // |dC_e| <= sum over the propagation steps of max |dAgg[recv e]|
      bwd_produce_rows(sh, cx, a.b.dC, e, valid, valid ? a.b.cBound[__ldg(a.recv + e)] : 0.f);
      bwd_hidden(sh, cx, meta[0], 0.f, nullptr, a.b.renc, a.b.dE, e, valid);
      bwd_hidden(sh, cx, meta[1], 0.f, nullptr, a.b.g2, a.b.dG2, e, valid);
      bwd_hidden(sh, cx, meta[2], 0.f, nullptr, a.b.g1, a.b.dG1, e, valid);
      // d rel_in: the first D_REL_IN columns of the last product
      epi_layer_out<false>(sh, cx, exp2i(-cx.e_in) * meta[3].x, [&](int c, int, float (&v)[HW]) {
        if (valid && c == 0) {
          float4* o = reinterpret_cast<float4*>(a.b.dRel + e * D_REL_IN + HW * cx.half);
          o[0] = make_float4(v[0], v[1], v[2], v[3]);
          o[1] = make_float4(v[4], v[5], v[6], v[7]);
          if (cx.half == 0) { o[2] = make_float4(v[8], v[9], v[10], v[11]); o[3] = make_float4(v[12], v[13], v[14], v[15]); }
        }
      });
      tile = slot_tile(k + 1, cx.slot, n_tiles);
    }
  }
  chain_teardown(tmem_base);
}

}  // namespace tc

// ------------------------------------------------------------------------------------------------ host
static int tc_bwd_attrs() {
  static thread_local DeviceOnce once;
  if (!once.need()) return AGX_OK;
  using namespace tc;
  AGX_CUDA_OK(cudaFuncSetAttribute(tc_bwd_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  AGX_CUDA_OK(cudaFuncSetAttribute(tc_bwd_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  AGX_CUDA_OK(cudaFuncSetAttribute(tc_bwd_node_encoder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  AGX_CUDA_OK(cudaFuncSetAttribute(tc_bwd_edge_encoder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  return AGX_OK;
}

static tc::BwdArgs bwd_args(const AgxGraphIn* g, const float* wts, size_t base_bytes, const TcBwdBuffers& b) {
  tc::BwdArgs a{};
  a.B = g->B; a.N = g->N; a.n_p = g->n_p;
  a.rows = (int64_t)g->B * g->N; a.E_cap = g->E_cap;
  a.row_ptr = g->row_ptr; a.recv = g->recv;
  a.blob = reinterpret_cast<const uint8_t*>(wts);
  a.L = tc::tc_layout(base_bytes);
  a.b = b;
  return a;
}
static int node_grid(const tc::BwdArgs& a) {
  const int64_t tiles = (a.rows + tc::TILE - 1) / tc::TILE;
  return (int)(tiles < num_sms() ? tiles : num_sms());
}

int tc_bwd_head(const AgxGraphIn* g, const float* wts, const PackedLayout& PL, size_t base_bytes, const TcBwdBuffers& b, const float* P_K,
                const float* d_pos, const float* d_motion, const float* motion, cudaStream_t st) {
  using namespace tc;
  if (int rc = tc_bwd_attrs()) return rc;
  BwdArgs a = bwd_args(g, wts, base_bytes, b);
  a.P_act = P_K; a.head_w = wts + PL.pred2_w; a.d_pos = d_pos; a.d_motion = d_motion; a.motion = motion;
  { ProfScope ps(AGX_KIND_OTHER, st);
    tc_bwd_head_kernel<<<node_grid(a), THREADS, SMEM_BYTES, st>>>(a); }
  AGX_LAUNCH_CHECK();
  return AGX_OK;
}

int tc_bwd_step(const AgxGraphIn* g, const float* wts, size_t base_bytes, const TcBwdBuffers& b, const float* P_k, cudaStream_t st) {
  using namespace tc;
  if (int rc = tc_bwd_attrs()) return rc;
  BwdArgs a = bwd_args(g, wts, base_bytes, b);
  a.P_act = P_k;
  { ProfScope ps(AGX_KIND_OTHER, st);
    tc_bwd_step_kernel<<<node_grid(a), THREADS, SMEM_BYTES, st>>>(a); }
  AGX_LAUNCH_CHECK();
  return AGX_OK;
}

int tc_bwd_node_encoder(const AgxGraphIn* g, const float* wts, size_t base_bytes, const TcBwdBuffers& b, cudaStream_t st) {
  using namespace tc;
  if (int rc = tc_bwd_attrs()) return rc;
  BwdArgs a = bwd_args(g, wts, base_bytes, b);
  { ProfScope ps(AGX_KIND_OTHER, st);
    tc_bwd_node_encoder_kernel<<<node_grid(a), THREADS, SMEM_BYTES, st>>>(a); }
  AGX_LAUNCH_CHECK();
  return AGX_OK;
}

int tc_bwd_edge_encoder(const AgxGraphIn* g, const float* wts, size_t base_bytes, const TcBwdBuffers& b, cudaStream_t st) {
  using namespace tc;
  if (g->E_cap <= 0) return AGX_OK;
  if (int rc = tc_bwd_attrs()) return rc;
  BwdArgs a = bwd_args(g, wts, base_bytes, b);
  const int64_t tiles = (g->E_cap + TILE - 1) / TILE;
  { ProfScope ps(AGX_KIND_OTHER, st);
    tc_bwd_edge_encoder_kernel<<<(int)(tiles < num_sms() ? tiles : num_sms()), THREADS, SMEM_BYTES, st>>>(a); }
  AGX_LAUNCH_CHECK();
  return AGX_OK;
}

}  // namespace agx
