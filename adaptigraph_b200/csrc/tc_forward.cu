// Tensor-core (tcgen05) kernels of the forward pass: same stages and buffers as forward.cu's fp32
// kernels (node_encoder, edge_encoder, node_update, node_update_head), with the dense 160x160 layers
// evaluated as split-fp16 MMAs (tc_chain.cuh).  Reference arithmetic: dynamics/gnn/model.py:129-313.
#include "common.cuh"
#include "tc_chain.cuh"
#include "tc_forward.cuh"
#include "tc_backward.cuh"

namespace agx {
namespace tc {

constexpr float MOTION_CLAMP = 100.f;  // model.py:85

// Per-layer MMA budget (tc_chain.cuh: LayerStep::nmma).  Measured with tests/bench/precision_study.py (10-step cloth rollout against
// float64, relations rebuilt every step; profiles/r02_precision_study.txt): the relation chain tolerates fp16-rounded activations
// (every relation's error is independent and seven of them are averaged into a particle: 3.4e-6 RMSE for all three layers together),
// the particle chains do not (each single layer costs 0.7-1.5e-5: their errors enter the residual stream directly).
#ifndef AGX_MMA_EDGE
#define AGX_MMA_EDGE 2      // relation_encoder.model.2 / .4 and the relation part of relation_propagator
#endif
// Going below two fp16 MMAs does not pay any more: with the residual term Ahi * Wlo issued as ONE fp8 MMA (kind::f8f6f4, e5m2 copy of
// A in tensor memory x e4m3 image of Wlo: 1.5 MMA units, same 5e-6 rollout RMSE, all parity tests green) the relation encoder went
// from 0.466 to 0.460 ms (r02k) -- at two MMAs per product the chain is no longer bound by the MMA count.  Not adopted
// (profiles/r02_experiment_fp8_residual.patch).  Neither is giving the two column parts of a layer separate accumulator columns
// (one accumulator hand-over per layer instead of two): 0.477 against 0.479 ms (r02i).
#ifndef AGX_MMA_NODE
#define AGX_MMA_NODE 3      // every particle-side layer
#endif

// ------------------------------------------------------------------------------------ weight images
struct PackSpec {
  const float* W; const float* bias; int ld, col0, K, F;   // source: W[n*ld + col0 + k], n < F, k < K
  int transpose;                                            // 1: source W[k*ld + col0 + n] instead (image of the transposed matrix)
};
struct PackArgs {
  PackSpec spec[T_NUM];
  uint8_t* blob;
  TcLayout L;
};

__global__ void __launch_bounds__(256) pack_tc_kernel(const PackArgs a) {
  __shared__ float red_w[8], red_s[8];
  __shared__ float sw_s;
  const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const PackSpec s = a.spec[t];
  const int kpad = tc_kpad(t);
  // the bias occupies the first padding column (k = K): it is multiplied by the constant 1 every A row carries there
  float wmax = 0.f, smax = 0.f;
  auto src = [&](int n, int k) { return s.transpose ? s.W[(size_t)k * s.ld + s.col0 + n] : s.W[(size_t)n * s.ld + s.col0 + k]; };
  for (int n = tid; n < s.F; n += 256) {
    float rs = 0.f;
    for (int k = 0; k < s.K; ++k) {
      const float w = fabsf(src(n, k));
      wmax = fmaxf(wmax, w);
      rs += w;
    }
    if (s.bias) { const float b = fabsf(s.bias[n]); wmax = fmaxf(wmax, b); rs += b; }
    smax = fmaxf(smax, rs);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    smax = fmaxf(smax, __shfl_xor_sync(0xffffffffu, smax, o));
  }
  if (lane == 0) { red_w[warp] = wmax; red_s[warp] = smax; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 8; ++w) { red_w[0] = fmaxf(red_w[0], red_w[w]); red_s[0] = fmaxf(red_s[0], red_s[w]); }
    const int sw = scale_exp(red_w[0]);
    sw_s = exp2i(sw);
    float4* meta = reinterpret_cast<float4*>(a.blob + a.L.meta);
    // {2^-sw, max_n (sum_k |W_nk| + |b_n|) padded for fp32 summation slack, -, -}
    meta[t] = make_float4(exp2i(-sw), red_s[0] * 1.0001f, 0.f, 0.f);
  }
  __syncthreads();
  const float sc = sw_s;
  uint8_t* hi_img = a.blob + a.L.img[t];
  uint8_t* lo_img = hi_img + (size_t)FP * kpad * 2;
  for (int i = tid; i < FP * kpad; i += 256) {
    const int n = i / kpad, k = i - n * kpad;
    float v = 0.f;
    if (n < s.F) {
      if (k < s.K) v = src(n, k) * sc;
      else if (k == s.K && s.bias) v = s.bias[n] * sc;
    }
    const __half h = __float2half_rn(v);
    const __half l = __float2half_rn(v - __half2float(h));
    const uint32_t off = img_offset(n, k, kpad);
    *reinterpret_cast<__half*>(hi_img + off) = h;
    *reinterpret_cast<__half*>(lo_img + off) = l;
  }
}

// ------------------------------------------------------------------------------------ common kernel prologue / epilogue
#ifdef AGX_TC_STAGGER
// experiment: CTA i starts (i % 4) * g_tc_stagger clocks late so that the SMs' memory phases do not line up
__device__ int g_tc_stagger = 0;
#endif
template <int NL>
__device__ __forceinline__ uint32_t chain_setup(Shared& sh, uint8_t* smem_raw, const LayerStep (&prog)[NL], const uint8_t* blob,
                                                const TcLayout& L, float4 (&meta)[NL]) {
  sh = carve_shared(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(sh.bar_wsmall, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&sh.bar_wfull[i], 1); mbar_init(&sh.bar_wempty[i], NSLOT); }
    for (int i = 0; i < NSLOT; ++i) {
      mbar_init(&sh.bar_in[i], SLOT_WARPS); mbar_init(&sh.bar_a[i], SLOT_WARPS);
      mbar_init(&sh.bar_accfull[i], 1); mbar_init(&sh.bar_accempty[i], SLOT_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == MMA_WARP0) tmem_alloc(sh.tmem_ptr, TMEM_COLS);
  pdl_wait();      // nothing above touches global memory: barrier init and the tensor-memory allocation overlap the previous kernel's tail
  pdl_trigger();
  const float4* m = reinterpret_cast<const float4*>(blob + L.meta);
#pragma unroll
  for (int l = 0; l < NL; ++l) meta[l] = m[prog[l].layer];
#ifdef AGX_TC_STAGGER
  if (tid == 0 && g_tc_stagger > 0) {
    const long long t0 = clock64(), d = (long long)(blockIdx.x & 3) * g_tc_stagger;
    while (clock64() - t0 < d) __nanosleep(200);
  }
#endif
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return *sh.tmem_ptr;
}

__device__ __forceinline__ void chain_teardown(uint32_t tmem_base) {
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == MMA_WARP0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

__device__ __forceinline__ EpiCtx make_ctx(uint32_t tmem_base) {
  EpiCtx cx;
  cx.lane = threadIdx.x & 31;
  cx.warp = threadIdx.x >> 5;
  cx.slot = cx.warp / SLOT_WARPS;
  const int w = cx.warp % SLOT_WARPS;
  cx.row = (w & 3) * 32 + cx.lane;
  cx.half = w >> 2;
  cx.tslot = tmem_base + cx.slot * SLOT_COLS + ((uint32_t)((w & 3) * 32) << 16);
  cx.acc_parity = 0;
  cx.e_in = 0;
  cx.bound_in = 1.f;
#ifdef AGX_TC_TIMELINE
  cx.tl_region = -1;
  cx.tl_i = 0;
#endif
  return cx;
}

__device__ __forceinline__ float max16(const float (&v)[HW], float m) {   // max |v|
#pragma unroll
  for (int i = 0; i < HW; ++i) m = fmaxf(m, fabsf(v[i]));
  return m;
}
// this thread's 16 columns [col0, col0 + 16) of blocked row `row`: the narrow last piece holds only columns 144..151
__device__ __forceinline__ void blk_store16(float* base, int64_t row, int col0, const float (&v)[HW]) {
  float* p = base + blk_off(row, col0);
  const float a[8] = {v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]};
  stg256(p, a);
  if (col0 < BLK_LAST) {
    const float b[8] = {v[8], v[9], v[10], v[11], v[12], v[13], v[14], v[15]};
    stg256(p + 8, b);
  }
}
__device__ __forceinline__ void blk_add16(const float* base, int64_t row, int col0, float (&v)[HW]) {   // v += stored columns
  const float* p = base + blk_off(row, col0);
  float t[8];
  ldg256(p, t);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] += t[i];
  if (col0 < BLK_LAST) {
    ldg256(p + 8, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[8 + i] += t[i];
  }
}
// Row-major variant of blk_store16 (stride FP).  Measured for Qr / Qs on cloth-2k x 128 (r02b): the aggregate's gather gets five
// whole lines per sender instead of ten half-used ones (-0.014 ms per launch), but the chains' thread-per-row stores then touch 32
// lines per instruction instead of 16 and node_encoder / node_update lose 0.033 ms each: a net loss, so every intermediate stays
// blocked.
__device__ __forceinline__ void row_store16(float* base, int64_t row, int col0, const float (&v)[HW]) {
  float* p = base + row * FP + col0;
  const float a[8] = {v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]};
  const float b[8] = {v[8], v[9], v[10], v[11], v[12], v[13], v[14], v[15]};
  stg256(p, a);
  stg256(p + 8, b);
}
struct NoExtra { __device__ void operator()(int, int, float (&)[HW]) const {} };
struct NoSide { __device__ void operator()(int, int, const float (&)[HW]) const {} };

// plain hidden layer (bias inside the MMA): relu(acc) -> the slot's next A; the row bound is propagated, no row maximum needed
template <bool LO = true>
__device__ __forceinline__ void epi_hidden(const Shared& sh, EpiCtx& cx, const float4 meta_l) {
  const float bound_next = fmaxf(cx.bound_in * meta_l.y, 1.f);
  const int e_next = scale_exp(bound_next);
  epi_layer_plain<LO>(sh, cx, exp2i(-cx.e_in) * meta_l.x, exp2i(e_next));
  cx.bound_in = bound_next;
  cx.e_in = e_next;
}

// same layer, with the fp32 activations handed to `side` (training: the backward needs every ReLU output)
template <bool LO = true, class Side>
__device__ __forceinline__ void epi_hidden_side(const Shared& sh, EpiCtx& cx, const float4 meta_l, Side side) {
  const float bound_next = fmaxf(cx.bound_in * meta_l.y, 1.f);
  const int e_next = scale_exp(bound_next);
  epi_layer_to_a<LO>(sh, cx, exp2i(-cx.e_in) * meta_l.x, exp2i(e_next), NoExtra{}, side);
  cx.bound_in = bound_next;
  cx.e_in = e_next;
}

// acc (bias inside the MMA) -> fp32 rows in HBM (A is left untouched); returns this thread's partial max |v|
template <bool ROW_MAJOR = false, class AllRead = NoHook>
__device__ __forceinline__ float epi_store_rows(const Shared& sh, EpiCtx& cx, const float4 meta_l, float* out, int64_t grow, bool valid,
                                                AllRead all_read = AllRead{}) {
  const float unscale = exp2i(-cx.e_in) * meta_l.x;
  float mx = 0.f;
  epi_layer_out<false>(sh, cx, unscale, [&](int, int col0, float (&v)[HW]) {
    mx = max16(v, mx);
    if (valid) {
      if (ROW_MAJOR) row_store16(out, grow, col0, v);
      else blk_store16(out, grow, col0, v);
    }
  }, all_read);
  return mx;
}

// k-th tile of this CTA's slot, or -1
__device__ __forceinline__ int slot_tile(int k_slot, int slot, int n_tiles) {
  const int t = (int)blockIdx.x + (k_slot * NSLOT + slot) * (int)gridDim.x;
  return t < n_tiles ? t : -1;
}

// ------------------------------------------------------------------------------------ C in 16-bit block fixed point
// AGX_PREC_TC_MIXED stores the per-relation term C = W_rel * renc + b (written once, read by every propagation step: 47 % of the
// bytes a model step moves when kept in fp32) as 16-bit fixed point with one power-of-two scale per (relation, 16-column piece):
// a row is C16_ROW = 320 bytes = 152 offset-binary uint16 (u = rint(c * 2^e) + 32768, |c * 2^e| <= 2^14) followed by 16 bytes of
// piece exponents (int8 e; byte 8 * (p & 1) + (p >> 1) belongs to piece p: the two column halves of an epilogue slot each own the
// pieces of one parity and write their five exponents with one 8-byte store).  Rows are plain row-major, so the CSR-ordered
// relations of consecutive receivers are one contiguous range: the aggregate stages them with one bulk (TMA) copy per receiver.
// Error: <= 2^-15 of the piece maximum per element (rollout RMSE 7e-7 on its own, tests/bench/precision_study.py).
constexpr int C16_ROW = 320;                // bytes per relation
constexpr int C16_EXP_OFF = 2 * BLK_COLS;   // 304: byte offset of the exponent block
constexpr float C16_MAGIC = 8421376.f;      // 2^23 + 32768: float bits of (v + MAGIC) carry rint(v) + 32768 in their low 16 bits

// this thread's 16 columns [col0, col0 + 16) of relation row `row` -> C16; returns the piece exponent
__device__ __forceinline__ int c16_store16(uint8_t* base, int64_t row, int col0, const float (&v)[HW]) {
  const float mx = max16(v, 0.f);
  const int e = scale_exp(mx);
  const float sc = exp2i(e);
  uint32_t w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t a = __float_as_uint(fmaf(v[2 * i], sc, C16_MAGIC)), b = __float_as_uint(fmaf(v[2 * i + 1], sc, C16_MAGIC));
    w[i] = __byte_perm(a, b, 0x5410);       // {a.lo16, b.lo16}
  }
  uint8_t* p = base + row * C16_ROW + 2 * col0;
#ifdef AGX_ABLATE_C16_STORE   // profiling aid: the relation term is computed but not written
  asm volatile("" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
  return e;
#endif
  if (col0 < BLK_LAST) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]),
                 "r"(w[5]), "r"(w[6]), "r"(w[7])
                 : "memory");
  } else {                                   // narrow last piece: columns 144..151
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  return e;
}

// ------------------------------------------------------------------------------------ edge encoder
struct EdgeArgs {
  const int32_t* row_ptr; const int32_t* send; const int32_t* recv;
  int64_t rows; int N; int64_t E_cap;
  const float* nfeat;
  const uint8_t* blob; TcLayout L;
  float* C;          // fp32 blocked rows (AGX_PREC_TC_F16X3) or C16 rows (AGX_PREC_TC_MIXED)
  // training (SAVE): row-major fp32 copies of what the backward reads -- the 17 relation inputs ([E][D_REL_IN]) and the three
  // ReLU outputs of the relation encoder ([E][FP])
  float* sv_rel_in; float* sv_g1; float* sv_g2; float* sv_renc;
};

// 17 relation inputs (model.py:224-253) of the K=32 first layer; half 0 builds inputs 0..15, half 1 input 16 (+ zero padding)
// r, s: flattened receiver / sender of the relation (r < 0: no relation in this row of the tile)
__device__ __forceinline__ void edge_inputs(const EdgeArgs& a, int r, int s, int half, float (&v)[HW]) {
#pragma unroll
  for (int i = 0; i < HW; ++i) v[i] = 0.f;
  if (r >= 0) {
    const float4* fr = reinterpret_cast<const float4*>(a.nfeat + (size_t)r * NFEAT);
    const float4* fs = reinterpret_cast<const float4*>(a.nfeat + (size_t)s * NFEAT);
    if (half == 0) {   // [attr_r, attr_s, |group_r - group_s|, hist diff 0..10]
      const float4 r3 = fr[3], s3 = fs[3], r0 = fr[0], s0 = fs[0], r1 = fr[1], s1 = fs[1], r2 = fr[2], s2 = fs[2];
      v[0] = r3.x; v[1] = r3.y; v[2] = s3.x; v[3] = s3.y; v[4] = fabsf(r3.z - s3.z);
      v[5] = r0.x - s0.x; v[6] = r0.y - s0.y; v[7] = r0.z - s0.z; v[8] = r0.w - s0.w;
      v[9] = r1.x - s1.x; v[10] = r1.y - s1.y; v[11] = r1.z - s1.z; v[12] = r1.w - s1.w;
      v[13] = r2.x - s2.x; v[14] = r2.y - s2.y; v[15] = r2.z - s2.z;
    } else {           // hist diff 11
      v[0] = fr[2].w - fs[2].w;
    }
  }
}

template <bool MIXED, bool SAVE = false>
__global__ void __launch_bounds__(THREADS, 1) tc_edge_encoder_kernel(const EdgeArgs a) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int NM = MIXED ? AGX_MMA_EDGE : 3;
  constexpr LayerStep prog[4] = {{T_RENC0, 2, IN_PRODUCER, 3}, {T_RENC2, 10, IN_EPILOGUE, NM}, {T_RENC4, 10, IN_EPILOGUE, NM},
                                 {T_RP_REL, 10, IN_EPILOGUE, NM}};
  Shared sh;
  float4 meta[4];
  const uint32_t tmem_base = chain_setup(sh, smem_raw, prog, a.blob, a.L, meta);
  const int64_t E = min((int64_t)a.row_ptr[a.rows], a.E_cap);
  const int n_tiles = (int)((E + TILE - 1) / TILE);
  const int warp = threadIdx.x >> 5;
  if (warp >= EPI_WARPS) {
    reg_dealloc_other();
    if (warp == LOAD_WARP) loader_role(sh, prog, a.blob, a.L, n_tiles, 4);
    else if (warp < LOAD_WARP) mma_role(sh, prog, warp - MMA_WARP0, tmem_base, n_tiles, 4);
  } else {
    reg_alloc_epilogue();
    EpiCtx cx = make_ctx(tmem_base);
#ifdef AGX_TC_TIMELINE
    if ((4) == AGX_TC_TIMELINE && cx.lane == 0 && (cx.warp == 0 || cx.warp == 4)) cx.tl_region = 3 + cx.warp / 4;
#endif
    // The input of a tile is gathered, scaled and split into registers while the previous tile's last layer is still in the
    // tensor pipe, and committed to the slot's A the moment that layer's MMAs have read it.
    uint32_t in_hi[8], in_lo[8];
    int in_exp = 0;
    float in_bound = 1.f;
    // The relation's endpoints are fetched a whole tile ahead (two dependent index loads) and their history records pulled
    // towards L2 / L1 as soon as they are known, so that produce() below waits for one short round trip instead of three long ones
    // (r02l timeline: the gather sat 3000 cycles in front of the last layer's epilogue).
    int nr = -1, ns = -1;
    auto fetch_endpoints = [&](int tile) {
      const int64_t e = (int64_t)tile * TILE + cx.row;
      nr = ns = -1;
      if (e < E) { nr = __ldg(a.recv + e); ns = __ldg(a.send + e); }
    };
    auto touch_records = [&]() {
      if (nr >= 0) {
        ns += (nr / a.N) * a.N;
        asm volatile("prefetch.global.L1 [%0];" ::"l"(a.nfeat + (size_t)nr * NFEAT));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(a.nfeat + (size_t)ns * NFEAT));
      }
    };
    auto produce = [&](int tile_of) {
      float vin[HW];
      edge_inputs(a, nr, ns, cx.half, vin);
      if (SAVE) {
        const int64_t e = (int64_t)tile_of * TILE + cx.row;
        if (e < E) {
          float4* o = reinterpret_cast<float4*>(a.sv_rel_in + e * D_REL_IN + HW * cx.half);
          o[0] = make_float4(vin[0], vin[1], vin[2], vin[3]);
          o[1] = make_float4(vin[4], vin[5], vin[6], vin[7]);
          if (cx.half == 0) { o[2] = make_float4(vin[8], vin[9], vin[10], vin[11]); o[3] = make_float4(vin[12], vin[13], vin[14], vin[15]); }
        }
      }
      in_bound = fmaxf(epi_exchange<true>(sh, cx, max16(vin, 0.f)), 1.f);
      if (cx.half == 1) vin[1] = 1.f;   // input 17: the constant that multiplies the bias column of relation_encoder.model.0
      in_exp = scale_exp(in_bound);
      split16(vin, exp2i(in_exp), in_hi, in_lo);
    };
    auto commit = [&]() {
      epi_store_packed(cx, 0, in_hi, in_lo);
      epi_signal(cx, &sh.bar_in[cx.slot]);
    };
    int tile = slot_tile(0, cx.slot, n_tiles);
    if (tile >= 0) { fetch_endpoints(tile); touch_records(); produce(tile); commit(); }
    for (int k = 0; tile >= 0; ++k) {
      const int64_t e = (int64_t)tile * TILE + cx.row;
      AGX_STAMP_EPI(cx, 30);
      cx.e_in = in_exp;
      cx.bound_in = in_bound;
      const int next = slot_tile(k + 1, cx.slot, n_tiles);
      if (next >= 0) fetch_endpoints(next);
      if (SAVE) {
        epi_hidden_side<true>(sh, cx, meta[0], [&](int, int col0, const float (&v)[HW]) { if (e < E) row_store16(a.sv_g1, e, col0, v); });
        if (next >= 0) touch_records();
        epi_hidden_side<true>(sh, cx, meta[1], [&](int, int col0, const float (&v)[HW]) { if (e < E) row_store16(a.sv_g2, e, col0, v); });
        epi_hidden_side<true>(sh, cx, meta[2], [&](int, int col0, const float (&v)[HW]) { if (e < E) row_store16(a.sv_renc, e, col0, v); });
      } else {
        epi_hidden<a_needs_lo(prog, 1)>(sh, cx, meta[0]);
        if (next >= 0) touch_records();
        epi_hidden<a_needs_lo(prog, 2)>(sh, cx, meta[1]);
        epi_hidden<a_needs_lo(prog, 3)>(sh, cx, meta[2]);
      }
      AGX_STAMP_EPI(cx, 31);
      if (next >= 0) produce(next);
      if (!MIXED) {
        epi_store_rows(sh, cx, meta[3], a.C, e, e < E, [&]() { if (next >= 0) commit(); });
      } else {
        uint8_t* c16 = reinterpret_cast<uint8_t*>(a.C);
        uint64_t exps = 0;                 // byte c = exponent of this thread's piece of chunk c
        epi_layer_out<false>(sh, cx, exp2i(-cx.e_in) * meta[3].x, [&](int c, int col0, float (&v)[HW]) {
          if (e < E) exps |= (uint64_t)(uint8_t)(int8_t)c16_store16(c16, e, col0, v) << (8 * c);
        }, [&]() { if (next >= 0) commit(); });
        if (e < E) *reinterpret_cast<uint64_t*>(c16 + e * C16_ROW + C16_EXP_OFF + 8 * cx.half) = exps;
      }
      tile = next;
    }
  }
  chain_teardown(tmem_base);
}

// ------------------------------------------------------------------------------------ node encoder
struct NodeArgs {
  const float* state; const float* attrs; const float* action; const float* p_instance; const float* physics;
  int B, N, n_p;
  const uint8_t* blob; TcLayout L;
  float* nfeat; float* P; float* A; float* Qr; float* Qs; float* rowmaxP; float* rowmaxA;
  float* S0;   // A_n + particle_encode: the residual input of propagation step 0 in one stream instead of two
  // training (SAVE): row-major fp32 copies for the backward -- the node inputs ([rows][D_NODE_IN]) and the encoder's ReLU outputs
  float* sv_p_in; float* sv_h1; float* sv_h2; float* sv_penc;
};

// node inputs (model.py:168-195) + the nfeat record; only half 0 carries data (6 real inputs of the K=16 layer)
__device__ __forceinline__ void node_inputs(const NodeArgs& a, int64_t r, int64_t rows, int half, float (&in)[HW]) {
#pragma unroll
  for (int i = 0; i < HW; ++i) in[i] = 0.f;
  if (r >= 0 && r < rows && half == 0) {
    const int b = (int)(r / a.N), n = (int)(r - (int64_t)b * a.N);
    float s[H_FIX][3];
#pragma unroll
    for (int h = 0; h < H_FIX; ++h) {
      const float* p = a.state + (((size_t)b * H_FIX + h) * a.N + n) * 3;
      s[h][0] = p[0]; s[h][1] = p[1]; s[h][2] = p[2];
    }
    const float a0 = a.attrs[r * 2 + 0], a1 = a.attrs[r * 2 + 1];
    const float grp = n < a.n_p ? a.p_instance[(size_t)b * a.n_p + n] : 0.f;
    float4* nf = reinterpret_cast<float4*>(a.nfeat + r * NFEAT);      // model.py:155-165 history record
    nf[0] = make_float4(s[1][0] - s[0][0], s[1][1] - s[0][1], s[1][2] - s[0][2], s[2][0] - s[1][0]);
    nf[1] = make_float4(s[2][1] - s[1][1], s[2][2] - s[1][2], s[3][0] - s[2][0], s[3][1] - s[2][1]);
    nf[2] = make_float4(s[3][2] - s[2][2], s[3][0], s[3][1], s[3][2]);
    nf[3] = make_float4(a0, a1, grp, 0.f);
    in[0] = a0; in[1] = a1;
    in[2] = n < a.n_p ? a.physics[b] : 0.f;                               // model.py:186-189
    in[3] = a.action[r * 3 + 0]; in[4] = a.action[r * 3 + 1]; in[5] = a.action[r * 3 + 2];
  }
}

template <bool SAVE = false>
__global__ void __launch_bounds__(THREADS, 1) tc_node_encoder_kernel(const NodeArgs a) {
  extern __shared__ uint8_t smem_raw[];
  constexpr LayerStep prog[6] = {{T_PENC0, 1, IN_PRODUCER, 3}, {T_PENC2, 10, IN_EPILOGUE, AGX_MMA_NODE}, {T_PENC4, 10, IN_EPILOGUE, AGX_MMA_NODE},
                                 {T_PP_ENC, 10, IN_EPILOGUE, AGX_MMA_NODE}, {T_RP_RECV, 10, IN_SAME, AGX_MMA_NODE}, {T_RP_SEND, 10, IN_SAME, AGX_MMA_NODE}};
  Shared sh;
  float4 meta[6];
  const uint32_t tmem_base = chain_setup(sh, smem_raw, prog, a.blob, a.L, meta);
  const int64_t rows = (int64_t)a.B * a.N;
  const int n_tiles = (int)((rows + TILE - 1) / TILE);
  const int warp = threadIdx.x >> 5;
  if (warp >= EPI_WARPS) {
    reg_dealloc_other();
    if (warp == LOAD_WARP) loader_role(sh, prog, a.blob, a.L, n_tiles, 6);
    else if (warp < LOAD_WARP) mma_role(sh, prog, warp - MMA_WARP0, tmem_base, n_tiles, 6);
  } else {
    reg_alloc_epilogue();
    EpiCtx cx = make_ctx(tmem_base);
#ifdef AGX_TC_TIMELINE
    if ((6) == AGX_TC_TIMELINE && cx.lane == 0 && (cx.warp == 0 || cx.warp == 4)) cx.tl_region = 3 + cx.warp / 4;
#endif
    int tile = slot_tile(0, cx.slot, n_tiles);
    for (int k = 0; tile >= 0; ++k) {
      const int64_t r = (int64_t)tile * TILE + cx.row;
      const bool valid = r < rows;
      float in[HW];
      node_inputs(a, r, rows, cx.half, in);
      const float mx = fmaxf(epi_exchange<true>(sh, cx, max16(in, 0.f)), 1.f);
      if (SAVE && valid && cx.half == 0) {
        float4* o = reinterpret_cast<float4*>(a.sv_p_in + r * D_NODE_IN);
        o[0] = make_float4(in[0], in[1], in[2], in[3]);
        o[1] = make_float4(in[4], in[5], 0.f, 0.f);
      }
      if (cx.half == 0) in[6] = 1.f;      // input 6: the constant that multiplies the bias column of particle_encoder.model.0
      cx.e_in = scale_exp(mx);
      cx.bound_in = mx;
      if (cx.half == 0) epi_store_a(cx, 0, in, exp2i(cx.e_in));   // the K=16 layer reads A columns 0..7 only
      epi_signal(cx, &sh.bar_in[cx.slot]);

      if (SAVE) {
        epi_hidden_side<true>(sh, cx, meta[0], [&](int, int col0, const float (&v)[HW]) { if (valid) row_store16(a.sv_h1, r, col0, v); });
        epi_hidden_side<true>(sh, cx, meta[1], [&](int, int col0, const float (&v)[HW]) { if (valid) row_store16(a.sv_h2, r, col0, v); });
      } else {
        epi_hidden<a_needs_lo(prog, 1)>(sh, cx, meta[0]);
        epi_hidden<a_needs_lo(prog, 2)>(sh, cx, meta[1]);
      }
      {  // particle_encode = particle_effect_0 (model.py:268-269): next A and the fp32 P rows (with their row maximum)
        const float4 m = meta[2];
        const float bound_next = fmaxf(cx.bound_in * m.y, 1.f);
        const int e_next = scale_exp(bound_next);
        float pm = epi_layer_to_a<a_needs_lo(prog, 3)>(sh, cx, exp2i(-cx.e_in) * m.x, exp2i(e_next), NoExtra{}, [&](int, int col0, const float (&v)[HW]) {
          if (valid) {
            blk_store16(a.P, r, col0, v);
            if (SAVE) row_store16(a.sv_penc, r, col0, v);
          }
        });
        pm = epi_exchange<true>(sh, cx, pm);
        cx.bound_in = bound_next;
        cx.e_in = e_next;
        if (valid && cx.half == 0) a.rowmaxP[r] = pm;
      }
      float am = 0.f;                                              // A_n = W_enc*penc + b, and S0 = A_n + particle_encode
      epi_layer_out<false>(sh, cx, exp2i(-cx.e_in) * meta[3].x, [&](int, int col0, float (&v)[HW]) {
        am = max16(v, am);
        if (valid) {
          blk_store16(a.A, r, col0, v);
          blk_add16(a.P, r, col0, v);        // this thread's own store of a moment ago (L1 / L2 hit)
          blk_store16(a.S0, r, col0, v);
        }
      });
      am = epi_exchange<true>(sh, cx, am);
      if (valid && cx.half == 0) a.rowmaxA[r] = am;
      epi_store_rows<false>(sh, cx, meta[4], a.Qr, r, valid);
      epi_store_rows<false>(sh, cx, meta[5], a.Qs, r, valid);
      tile = slot_tile(k + 1, cx.slot, n_tiles);
    }
  }
  chain_teardown(tmem_base);
}

// The particle encoder's inputs are attrs, the physics parameter and the action (model.py:168-195; state_dim is 0): none of them
// changes between the steps of a rollout (forward_dynamics.py:156-197 keeps `action` fixed within a push), so its products --
// particle_encode = P, A_n, Qr, Qs of propagation step 0 -- are computed by the first model step only (agx_rollout) and every later
// step refreshes just the history record the relation inputs are gathered from.
__global__ void __launch_bounds__(256) nfeat_kernel(const NodeArgs a) {
  const int64_t rows = (int64_t)a.B * a.N;
  const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (r >= rows) return;
  float in[HW];
  node_inputs(a, r, rows, 0, in);
}

// ------------------------------------------------------------------------------------ node update / head
struct UpdArgs {
  int B, N, n_p;
  const uint32_t* agg_split;   // blocked; per (row, 16-column piece) 8 packed-fp16 hi words then 8 lo words (edge_aggregate, split output)
  const int32_t* agg_exp; const float* agg_max;   // per-row scale exponent / row maximum of agg
  const float* A; const float* P_in; float* P; float* Qr; float* Qs; const float* rowmaxP_in; float* rowmaxP; const float* rowmaxA;
  const float* S0;   // propagation step 0: A_n + P_in precomputed by the particle encoder (nullptr otherwise)
  const uint8_t* blob; TcLayout L;   // P_in / rowmaxP_in: the residual stream this step reads (the encoder's copy at pstep 0, else P / rowmaxP)
  const float* head_w;   // fp32 [3][FP] then 4 bias floats (head only)
  const float* state; float* pred_pos; int64_t pos_stride_b; float* pred_motion;
  // training (SAVE): row-major fp32 copies for the backward -- the new particle effects and, in the head, its two ReLU outputs
  float* sv_P; float* sv_u1; float* sv_u2;
};

// AGG32: the aggregate arrives as plain blocked fp32 rows (edge_aggregate_c16: that kernel is bound by instruction issue, this one
// by HBM with issue slots to spare, so the row maximum / scale / fp16 split of the aggregate moved here: same bits).
template <bool LAST, bool SAVE = false, bool AGG32 = false>
__global__ void __launch_bounds__(THREADS, 1) tc_node_update_kernel(const UpdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  constexpr LayerStep prog[3] = {{T_PP_AGG, 10, IN_PRODUCER, 3}, {LAST ? T_PRED0 : T_RP_RECV, 10, IN_EPILOGUE, AGX_MMA_NODE},
                                 {LAST ? T_PRED1 : T_RP_SEND, 10, LAST ? IN_EPILOGUE : IN_SAME, AGX_MMA_NODE}};
  Shared sh;
  float4 meta[3];
  const uint32_t tmem_base = chain_setup(sh, smem_raw, prog, a.blob, a.L, meta);
  if (LAST) {
    for (int i = threadIdx.x; i < 3 * FP + 4; i += THREADS) sh.head_w[i] = a.head_w[i];
    __syncthreads();
  }
  const int64_t rows = (int64_t)a.B * a.N;
  const int n_tiles = (int)((rows + TILE - 1) / TILE);
  const int warp = threadIdx.x >> 5;
  if (warp >= EPI_WARPS) {
    reg_dealloc_other();
    if (warp == LOAD_WARP) loader_role(sh, prog, a.blob, a.L, n_tiles, LAST ? 13 : 3, !LAST);
    else if (warp < LOAD_WARP) mma_role(sh, prog, warp - MMA_WARP0, tmem_base, n_tiles, LAST ? 13 : 3);
  } else {
    reg_alloc_epilogue();
    EpiCtx cx = make_ctx(tmem_base);
#ifdef AGX_TC_TIMELINE
    if ((LAST ? 13 : 3) == AGX_TC_TIMELINE && cx.lane == 0 && (cx.warp == 0 || cx.warp == 4)) cx.tl_region = 3 + cx.warp / 4;
#endif
    int tile = slot_tile(0, cx.slot, n_tiles);
    // AGG32: maximum of this thread's row of tile t over both column halves (agg >= 0: a sum of ReLUs)
    auto agg_row_max = [&](int t) {
      const int64_t rr = (int64_t)t * TILE + cx.row;
      float mx = 0.f;
      if (rr < rows) {
        const float* agg = reinterpret_cast<const float*>(a.agg_split);
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c) {
          const int col0 = 32 * c + HW * cx.half;
          const float* p = agg + blk_off(rr, col0);
          float t8[8];
          ldg256(p, t8);
#pragma unroll
          for (int i = 0; i < 8; ++i) mx = fmaxf(mx, t8[i]);
          if (col0 < BLK_LAST) {
            ldg256(p + 8, t8);
#pragma unroll
            for (int i = 0; i < 8; ++i) mx = fmaxf(mx, t8[i]);
          }
        }
      }
      return epi_exchange<true>(sh, cx, mx);
    };
    for (int k = 0; tile >= 0; ++k) {
      const int64_t r = (int64_t)tile * TILE + cx.row;
      const bool valid = r < rows;
      const int next = slot_tile(k + 1, cx.slot, n_tiles);
      AGX_STAMP_EPI(cx, 30);
      if (cx.warp % SLOT_WARPS == 0 && cx.lane == 0) {
        // the residual rows of this tile are read by the first layer's epilogue: start them towards L2 now
        if (a.S0) {
          bulk_prefetch_l2(a.S0 + (int64_t)tile * BLK_TILE, BLK_TILE * 4);   // (buffers are padded to whole tiles)
        } else {
          bulk_prefetch_l2(a.A + (int64_t)tile * BLK_TILE, BLK_TILE * 4);
          bulk_prefetch_l2(a.P_in + (int64_t)tile * BLK_TILE, BLK_TILE * 4);
        }
      }
      if (AGG32) {
        // ---- producer: fp32 rows -> row maximum (both column halves) -> exact power-of-two scale -> split into A.  Two passes over
        // the row, back to back: the second read comes from L2 (80 floats per thread would not stay in registers).  split16 rounds
        // exactly like edge_aggregate_split_kernel's own conversion, so both routes put the same bits into tensor memory.
        // Measured alternatives (cloth-2k x 128, ms per launch of this kernel; 0.180 with the pre-split input): these two passes
        // 0.194; the maximum taken a tile ahead, behind the first layer's MMAs: 0.246 (a tile period streams ~130 MB through L2, the
        // rows are gone when the split wants them); two partial maxima per row left by the aggregate: 0.183 here, +0.016 there.
        const float* agg = reinterpret_cast<const float*>(a.agg_split);
        const float mx = agg_row_max(tile);
        cx.e_in = scale_exp(mx);
        cx.bound_in = mx;                                // actual row maximum (W_agg has no bias: no constant column needed)
        const float sc = exp2i(cx.e_in);
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c) {
          const int col0 = 32 * c + HW * cx.half;
          float v[HW];
#pragma unroll
          for (int i = 0; i < HW; ++i) v[i] = 0.f;
          if (valid) {
            const float* p = agg + blk_off(r, col0);
            float t[8];
            ldg256(p, t);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = t[i];
            if (col0 < BLK_LAST) {
              ldg256(p + 8, t);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[8 + i] = t[i];
            }
          }
          epi_store_a(cx, c, v, sc);
        }
        epi_signal(cx, &sh.bar_in[cx.slot]);
      } else {
        // ---- producer: the aggregated relation effects arrive already scaled and split (edge_aggregate): copy to A
        {
#pragma unroll
          for (int c = 0; c < NCHUNK; ++c) {
            uint32_t hi[8] = {0, 0, 0, 0, 0, 0, 0, 0}, lo[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            if (valid) {   // 16 words per (row, piece): 8 hi then 8 lo; the narrow last piece: 4 hi then 4 lo
              const int col0 = 32 * c + HW * cx.half;
              const uint32_t* piece = a.agg_split + blk_off(r, col0);
              if (col0 < BLK_LAST) {
                ldg256(reinterpret_cast<const float*>(piece), reinterpret_cast<float(&)[8]>(hi));
                ldg256(reinterpret_cast<const float*>(piece + 8), reinterpret_cast<float(&)[8]>(lo));
              } else {
                uint32_t w[8];
                ldg256(reinterpret_cast<const float*>(piece), reinterpret_cast<float(&)[8]>(w));
#pragma unroll
                for (int i = 0; i < 4; ++i) { hi[i] = w[i]; lo[i] = w[4 + i]; }
              }
            }
            epi_store_packed(cx, c, hi, lo);
          }
          cx.e_in = valid ? a.agg_exp[r] : 0;
          cx.bound_in = valid ? a.agg_max[r] : 0.f;     // actual row maximum (W_agg has no bias: no constant column needed)
          epi_signal(cx, &sh.bar_in[cx.slot]);
        }
      }
      AGX_STAMP_EPI(cx, 31);
      {  // P <- relu((W_agg*agg + A_n) + P)   (model.py:36-40, :299-301)
        const float4 m = meta[0];
        const float extra_bound = valid ? a.rowmaxA[r] + a.rowmaxP_in[r] : 0.f;   // bound on |A_n + P| of this row
        const float bound_next = fmaxf(cx.bound_in * m.y + extra_bound, 1.f);
        const int e_next = scale_exp(bound_next);
        const float unscale = exp2i(-cx.e_in) * m.x;
        float pm = epi_layer_to_a<a_needs_lo(prog, 1)>(sh, cx, unscale, exp2i(e_next),
                                  [&](int, int col0, float (&v)[HW]) {
                                    if (valid) {
                                      if (a.S0) {
                                        blk_add16(a.S0, r, col0, v);
                                      } else {
                                        blk_add16(a.A, r, col0, v);
                                        blk_add16(a.P_in, r, col0, v);
                                      }
                                    }
                                  },
                                  [&](int, int col0, const float (&v)[HW]) {
                                    if (!LAST && valid) blk_store16(a.P, r, col0, v);
                                    if (SAVE && valid) row_store16(a.sv_P, r, col0, v);
                                  });
        pm = epi_exchange<true>(sh, cx, pm);
        cx.bound_in = fmaxf(pm, 1.f);                  // actual maximum of the new P row (tighter than the propagated bound)
        cx.e_in = e_next;
        if (!LAST && valid && cx.half == 0) a.rowmaxP[r] = pm;
      }
      if (next >= 0 && cx.warp % SLOT_WARPS == 0 && cx.lane == 0)   // the next tile's input rows: two layers of lead time
        bulk_prefetch_l2(a.agg_split + (int64_t)next * BLK_TILE, BLK_TILE * 4);
      if (!LAST) {
        epi_store_rows<false>(sh, cx, meta[1], a.Qr, r, valid);
        epi_store_rows<false>(sh, cx, meta[2], a.Qs, r, valid);
      } else {
        if (SAVE) epi_hidden_side<true>(sh, cx, meta[1], [&](int, int col0, const float (&v)[HW]) { if (valid) row_store16(a.sv_u1, r, col0, v); });
        else epi_hidden<a_needs_lo(prog, 2)>(sh, cx, meta[1]);
        // motion head (model.py:306-309): relu(linear_1) then the 3-row linear_2 as running dot products
        const float unscale = exp2i(-cx.e_in) * meta[2].x;
        float m0 = 0.f, m1 = 0.f, m2 = 0.f;
        epi_layer_out<true>(sh, cx, unscale, [&](int, int col0, float (&v)[HW]) {
          if (SAVE && valid) row_store16(a.sv_u2, r, col0, v);
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const float4 w0 = lds128(sh.head_w + col0 + 4 * h), w1 = lds128(sh.head_w + FP + col0 + 4 * h),
                         w2 = lds128(sh.head_w + 2 * FP + col0 + 4 * h);
            m0 = fmaf(v[4 * h + 3], w0.w, fmaf(v[4 * h + 2], w0.z, fmaf(v[4 * h + 1], w0.y, fmaf(v[4 * h], w0.x, m0))));
            m1 = fmaf(v[4 * h + 3], w1.w, fmaf(v[4 * h + 2], w1.z, fmaf(v[4 * h + 1], w1.y, fmaf(v[4 * h], w1.x, m1))));
            m2 = fmaf(v[4 * h + 3], w2.w, fmaf(v[4 * h + 2], w2.z, fmaf(v[4 * h + 1], w2.y, fmaf(v[4 * h], w2.x, m2))));
          }
        });
        m0 = epi_exchange<false>(sh, cx, m0);
        m1 = epi_exchange<false>(sh, cx, m1);
        m2 = epi_exchange<false>(sh, cx, m2);
        if (valid && cx.half == 0) {
          const int b = (int)(r / a.N), n = (int)(r - (int64_t)b * a.N);
          if (n < a.n_p) {
            m0 += lds32(sh.head_w + 3 * FP + 0); m1 += lds32(sh.head_w + 3 * FP + 1); m2 += lds32(sh.head_w + 3 * FP + 2);
            float* mo = a.pred_motion + ((size_t)b * a.n_p + n) * 3;
            mo[0] = m0; mo[1] = m1; mo[2] = m2;
            const float* cur = a.state + (((size_t)b * H_FIX + (H_FIX - 1)) * a.N + n) * 3;
            float* po = a.pred_pos + (size_t)b * a.pos_stride_b + (size_t)n * 3;
            po[0] = cur[0] + fminf(fmaxf(m0, -MOTION_CLAMP), MOTION_CLAMP);
            po[1] = cur[1] + fminf(fmaxf(m1, -MOTION_CLAMP), MOTION_CLAMP);
            po[2] = cur[2] + fminf(fmaxf(m2, -MOTION_CLAMP), MOTION_CLAMP);
          }
        }
      }
      tile = next;
    }
  }
  chain_teardown(tmem_base);
}

// ------------------------------------------------------------------------------------ single dense layer (training path)
// Y[m][0:160] (+)= [relu]( (X[m][0:160] (*) [mask[m][k] > 0]) * W + bias + add1[m] + add2[m] )  — the contract of train.cu's
// lin_kernel<160>, with the product on the tensor cores (same split-fp16 scheme, same 128-row tiles and two slots per CTA as the
// chains above; rows are plain row-major fp32 here because that is what the backward kernels read).  `layer` selects the weight
// image: a T_* image for a forward product (its bias column is switched off: the fp32 bias is added in the epilogue) or a TT_*
// image (transposed matrix) for the backward's dX = dY * W.
struct LinTcArgs {
  const float* X; int ldx;
  const float* mask; int ldm;
  const float* bias; const float* add1; const float* add2;
  float* Y; int ldy;
  int64_t M;
  int relu, accumulate, n_store, layer;
  const uint8_t* blob; TcLayout L;
};

__global__ void __launch_bounds__(THREADS, 1) tc_lin_kernel(const LinTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const LayerStep prog[1] = {{a.layer, 10, IN_PRODUCER, 3}};
  Shared sh;
  float4 meta[1];
  const uint32_t tmem_base = chain_setup(sh, smem_raw, prog, a.blob, a.L, meta);
  const int n_tiles = (int)((a.M + TILE - 1) / TILE);
  const int warp = threadIdx.x >> 5;
  if (warp >= EPI_WARPS) {
    reg_dealloc_other();
    if (warp == LOAD_WARP) loader_role(sh, prog, a.blob, a.L, n_tiles, 20);
    else if (warp < LOAD_WARP) mma_role(sh, prog, warp - MMA_WARP0, tmem_base, n_tiles, 20);
  } else {
    reg_alloc_epilogue();
    EpiCtx cx = make_ctx(tmem_base);
    int tile = slot_tile(0, cx.slot, n_tiles);
    for (int k = 0; tile >= 0; ++k) {
      const int64_t r = (int64_t)tile * TILE + cx.row;
      const bool valid = r < a.M;
      const float* xrow = a.X + r * a.ldx;
      const float* mrow = a.mask ? a.mask + r * a.ldm : nullptr;
      // this thread's 16 columns of chunk c, masked; columns >= 150 (padding, and the image's bias column) are forced to zero
      auto load_piece = [&](int c, float (&v)[HW]) {
        const int col0 = 32 * c + HW * cx.half;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
          if (valid) {
            x = *reinterpret_cast<const float4*>(xrow + col0 + 4 * q);
            if (mrow) {
              const float4 m = *reinterpret_cast<const float4*>(mrow + col0 + 4 * q);
              x.x = m.x > 0.f ? x.x : 0.f; x.y = m.y > 0.f ? x.y : 0.f; x.z = m.z > 0.f ? x.z : 0.f; x.w = m.w > 0.f ? x.w : 0.f;
            }
          }
          v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
        }
        if (c == NCHUNK - 1 && cx.half == 1) {
#pragma unroll
          for (int i = ONE_COL % HW; i < HW; ++i) v[i] = 0.f;
        }
      };
      // ---- producer, pass 1: row maximum -> exact power-of-two scale; pass 2: split into the slot's A (second read hits L1 / L2)
      float mx = 0.f;
#pragma unroll
      for (int c = 0; c < NCHUNK; ++c) {
        float v[HW];
        load_piece(c, v);
        mx = max16(v, mx);
      }
      mx = epi_exchange<true>(sh, cx, mx);
      cx.e_in = scale_exp(mx);
      cx.bound_in = mx;
      const float sc = exp2i(cx.e_in);
#pragma unroll
      for (int c = 0; c < NCHUNK; ++c) {
        float v[HW];
        load_piece(c, v);
        epi_store_a(cx, c, v, sc);
      }
      epi_signal(cx, &sh.bar_in[cx.slot]);
      // ---- epilogue
      const float unscale = exp2i(-cx.e_in) * meta[0].x;
      float* yrow = a.Y + r * a.ldy;
      epi_layer_out<false>(sh, cx, unscale, [&](int, int col0, float (&v)[HW]) {
        if (!valid) return;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int col = col0 + 4 * q;
          if (col >= a.n_store) continue;                     // n_store is a multiple of 4
          float4 o = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          if (a.bias) { const float4 b = *reinterpret_cast<const float4*>(a.bias + col); o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w; }
          if (a.add1) { const float4 b = *reinterpret_cast<const float4*>(a.add1 + r * FP + col); o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w; }
          if (a.add2) { const float4 b = *reinterpret_cast<const float4*>(a.add2 + r * FP + col); o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w; }
          if (a.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
          float4* y = reinterpret_cast<float4*>(yrow + col);
          if (a.accumulate) { const float4 p = *y; o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
          *y = o;
        }
      });
      tile = slot_tile(k + 1, cx.slot, n_tiles);
    }
  }
  chain_teardown(tmem_base);
}

// ------------------------------------------------------------------------------------ edge aggregate (split output)
// Same reduction as forward.cu's edge_aggregate_kernel, but the row leaves the kernel the way the update chain's
// tensor-memory A wants it: scaled by an exact per-row power of two and split into packed fp16 hi / lo columns.
constexpr int AGG_NODES = 8;
constexpr int AGG_THREADS = AGG_NODES * (FP / 4);  // 320

__device__ __forceinline__ float4 relu_add3(const float4 c, const float4 qr, const float4 qs) {
  return make_float4(fmaxf((c.x + qr.x) + qs.x, 0.f), fmaxf((c.y + qr.y) + qs.y, 0.f),
                     fmaxf((c.z + qr.z) + qs.z, 0.f), fmaxf((c.w + qr.w) + qs.w, 0.f));
}


// Persistent: CTA b handles the row groups b, b + grid, b + 2 grid, ... (AGG_NODES consecutive receivers each, 40 threads = one
// float4 column each per receiver), so the grid sweeps a contiguous window of rows and of the CSR-ordered C.  The index data is
// software-pipelined: a group's sender ids are copied into shared memory one group ahead (cp.async, so no register is tied up while
// they are in flight) and its row_ptr pair is loaded two groups ahead, so per group a thread waits for exactly one round trip -- the
// 2 * AGG_BATCH feature-row loads it has in flight.  Summation is in CSR order per
// column (deterministic, no atomics on data); the row maximum goes through a triple-buffered shared-memory slot.
#ifndef AGX_AGG_BATCH
#define AGX_AGG_BATCH 8
#define AGX_AGG_CTAS 2
#endif
constexpr int AGG_BATCH = AGX_AGG_BATCH;   // relations in flight per thread
constexpr int AGG_CTAS_PER_SM = AGX_AGG_CTAS;
// Measured on cloth-2k x 128 (B200, ms per launch): one CTA per 8 rows with 2 relations in flight 0.392; this kernel with the sender
// ids prefetched into registers and 64-bit indexing, (batch, CTAs/SM) = (8, 2) 0.349 [32-bit indexing 0.332, ids through cp.async 0.308], (4, 3) 0.386, (2, 4) 0.370, (6, 2) 0.436; visiting runs of 16 consecutive groups per CTA 0.387;
// a warp-per-row variant staging C through per-warp TMA rings 0.658 (150 instructions per relation: issue-latency bound);
// this kernel plus a bulk L2 prefetch of the next group's C / Qr rows 0.389 (the memory system is request-throughput bound on the
// 64-byte pieces of the blocked layout, not latency bound: more requests in flight only queue); sweeping the rows in reverse so that
// the most recently written C / Qr / Qs tiles are read first: no change (0.308 both ways); 8 columns per thread with 256-bit loads
// (19 threads per row, 4 relations in flight, half the load instructions per byte): 0.328-0.335 against 0.310-0.313 on the same box.
__global__ void __launch_bounds__(AGG_THREADS, AGG_CTAS_PER_SM) edge_aggregate_split_kernel(
    const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ send, int rows, int N, int E_cap,
    const float4* __restrict__ C, const float4* __restrict__ Qr, const float4* __restrict__ Qs, uint32_t* __restrict__ agg_split,
    int32_t* __restrict__ agg_exp, float* __restrict__ agg_max, float4* __restrict__ agg_f32 /* optional row-major copy (training) */) {
  __shared__ int smax[3][AGG_NODES];
  __shared__ __align__(16) int32_t ids[2][AGG_NODES][AGG_BATCH];   // first AGG_BATCH sender ids of every row of this / the next group
  const int slot = threadIdx.x / (FP / 4), j = threadIdx.x - slot * (FP / 4);
  if (threadIdx.x < 3 * AGG_NODES) (&smax[0][0])[threadIdx.x] = 0;
  pdl_wait();
  pdl_trigger();
  const int n_groups = (rows + AGG_NODES - 1) / AGG_NODES, G = gridDim.x;
  // float4 j of a row in the blocked layout (tc_chain.cuh: blk_off), as a 32-bit float4 index (the host checks that it fits):
  // (row / 128) * tile + piece * 128 * 4 + (row % 128) * (floats4 per piece row) + j % 4; the narrow last piece has two float4 per
  // row, so the threads j = 38, 39 of a row have nothing to load (their columns 152..159 are padding: they keep zeros)
  const bool has_cols = j < BLK_COLS / 4;
  const int jj = has_cols ? j : BLK_COLS / 4 - 1;      // the two idle threads shadow the row's last float4 (same address: no traffic)
  const bool narrow = jj >= BLK_LAST / 4;
  // index = row * jrow + (row / 128) * jtile + joff, all per-thread constants
  const uint32_t jrow = narrow ? (BLK_COLS - BLK_LAST) / 4 : BLK_W / 4;
  const uint32_t jtile = BLK_TILE / 4 - TILE * jrow;
  const uint32_t joff = (uint32_t)(jj >> 2) * (TILE * BLK_W / 4) + (jj & 3);
  auto at = [=](const float4* m, int row) { return __ldg(m + ((uint32_t)row * jrow + ((uint32_t)row >> 7) * jtile + joff)); };
  auto load_bounds = [&](int v, int& beg, int& end) {   // [beg, end) of this thread's row in group v (empty past the end)
    const int r = v * AGG_NODES + slot;
    beg = end = 0;
    if (v < n_groups && r < rows) { beg = min(__ldg(row_ptr + r), E_cap); end = min(__ldg(row_ptr + r + 1), E_cap); }
  };
  // threads j < AGG_BATCH of every row copy sender id j of [beg, end) straight into shared memory (cp.async: no register is tied up
  // while the id is in flight); slots past the end are never read
  auto stage_senders = [&](int buf, int beg, int end) {
    if (j < AGG_BATCH && beg + j < end)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(&ids[buf][slot][j])), "l"(send + beg + j) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int beg, end, beg1, end1;
  load_bounds(blockIdx.x, beg, end);
  load_bounds(blockIdx.x + G, beg1, end1);
  stage_senders(0, beg, end);
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  for (int it = 0, v = blockIdx.x; v < n_groups; v += G, ++it) {
    const int r = v * AGG_NODES + slot;
    const bool valid = r < rows;
    const int gb = valid ? (r / N) * N : 0;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 c[AGG_BATCH], q[AGG_BATCH];
    const int n0 = min(end - beg, AGG_BATCH);
    float4 qr = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) qr = at(Qr, r);
    if (n0 > 0) {
      const int32_t* my = ids[it & 1][slot];
#pragma unroll
      for (int u = 0; u < AGG_BATCH; ++u) {
        const int uu = min(u, n0 - 1);                    // past-the-end slots repeat the last relation (masked below)
        c[u] = at(C, beg + uu);
        q[u] = at(Qs, gb + my[uu]);
      }
    }
    // index data of the groups to come (their latency hides behind this group's feature rows)
    int beg2, end2;
    stage_senders((it + 1) & 1, beg1, end1);
    load_bounds(v + 2 * G, beg2, end2);
    if (n0 > 0) {
#pragma unroll
      for (int u = 0; u < AGG_BATCH; ++u) {
        if (u < n0) {
          const float4 w = relu_add3(c[u], qr, q[u]);
          acc.x += w.x; acc.y += w.y; acc.z += w.z; acc.w += w.w;
        }
      }
    }
    for (int e0 = beg + AGG_BATCH; e0 < end; e0 += AGG_BATCH) {   // rows with more than AGG_BATCH relations: further (unpipelined) batches
      const int n = min(end - e0, AGG_BATCH);
#pragma unroll
      for (int u = 0; u < AGG_BATCH; ++u) {
        const int uu = min(u, n - 1);
        c[u] = at(C, e0 + uu);
        q[u] = at(Qs, gb + __ldg(send + e0 + uu));
      }
#pragma unroll
      for (int u = 0; u < AGG_BATCH; ++u) {
        if (u < n) {
          const float4 w = relu_add3(c[u], qr, q[u]);
          acc.x += w.x; acc.y += w.y; acc.z += w.z; acc.w += w.w;
        }
      }
    }
    // agg >= 0 (sum of ReLUs): the int view of the floats orders like the floats
    int* mxs = smax[it % 3];
    if (valid) atomicMax(&mxs[slot], __float_as_int(fmaxf(fmaxf(acc.x, acc.y), fmaxf(acc.z, acc.w))));
    asm volatile("cp.async.wait_group 0;" ::: "memory");   // the next group's sender ids have landed ...
    __syncthreads();                                       // ... and are visible to the whole CTA; row maxima complete
    if (threadIdx.x < AGG_NODES) smax[(it + 2) % 3][threadIdx.x] = 0;   // free since the previous barrier; next used after the next one
    if (valid) {
      const float mx = __int_as_float(mxs[slot]);
      const int e = scale_exp(mx);
      const float sc = exp2i(e);
      const float2 s0 = __fmul2_rn(make_float2(acc.x, acc.y), make_float2(sc, sc)), s1f = __fmul2_rn(make_float2(acc.z, acc.w), make_float2(sc, sc));
      const __half2 h0 = __float22half2_rn(s0), h1 = __float22half2_rn(s1f);
      const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
      const __half2 l0 = __float22half2_rn(make_float2(s0.x - f0.x, s0.y - f0.y)), l1 = __float22half2_rn(make_float2(s1f.x - f1.x, s1f.y - f1.y));
      // the 16 words of (row, piece = 16 columns): 8 packed hi pairs then 8 packed lo pairs (narrow last piece: 4 then 4)
      if (has_cols) {
        uint32_t* piece = agg_split + blk_off(r, 16 * (j >> 2));
        const int lo_at = narrow ? (BLK_COLS - BLK_LAST) / 2 : BLK_W / 2;
        *reinterpret_cast<uint2*>(piece + 2 * (j & 3)) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
        *reinterpret_cast<uint2*>(piece + lo_at + 2 * (j & 3)) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
      }
      if (j == 0) { agg_exp[r] = e; agg_max[r] = mx; }
      if (agg_f32) agg_f32[(size_t)r * (FP / 4) + j] = acc;
    }
    beg = beg1; end = end1; beg1 = beg2; end1 = end2;
  }
}


// ------------------------------------------------------------------------------------ edge aggregate on C16 (AGX_PREC_TC_MIXED)
// Same reduction with C in the 16-bit block format above, the row leaving as plain blocked fp32 (the update chain scales and splits
// it, tc_node_update_kernel<.., AGG32>).  38 threads per receiver (4 columns each: 8 bytes of C16, one float4 of the blocked fp32
// Qr / Qs rows), 5 receivers per group, persistent CTAs as above.  C never goes through registers or L1 on its
// way in: the first A16_BATCH relations of a receiver are one contiguous run of C16 rows, which thread 0 of the receiver's slot
// brings into shared memory with ONE bulk copy of the TMA engine (cp.async.bulk + mbarrier transaction count), a whole group ahead
// of its use -- next to the sender ids (cp.async) and the row_ptr pairs (two groups ahead).  What a thread waits for per group is
// therefore only the Qs gather it has in flight (A16_BATCH float4).  A task is (receiver group, batch of A16_BATCH relations per
// receiver): receivers of higher degree (granular: up to 25) take further batches through the same pipeline.
// The inner loop is branch-free: all A16_BATCH staged slots are always evaluated and a 0 / 1 factor in the accumulating FMA drops
// the ones past the receiver's degree (they read whatever well-formed rows and sender ids an earlier group left in shared memory,
// which is zero-initialised), because on this kernel the instruction issue and the L1 data pipe, not DRAM, are the bound: r02a's
// ncu capture of the first version (per-relation branches, generic loads, 64-register build with spills) showed 120 issued warp
// instructions per relation, 39 % DRAM utilisation and the L1 data pipe at 73 %.
// Where the time went with the split output still in this kernel (cloth-2k x 128, 0.249 ms per launch; r02i / r02j, each line one
// thing removed): no Qs gather 0.190, no
// arithmetic 0.171, no output 0.226, no bulk copies 0.232, no shared-memory reads of C 0.237 -- no single limiter; issue slots
// (87 warp instructions per relation, of which 35 are the per-relation loop) and the gather latency share it.  Tried and not
// adopted (profiles/r02_experiments_not_adopted.patch): issuing the gather of the NEXT task before computing the current one
// (sender ids staged two tasks ahead; 120 registers, 456 threads per SM): 0.273 ms; CTA shapes 8 receivers x 2 CTAs per SM 0.262,
// 2 x 8 0.26, 4 x 3 0.26 against 4 x 4 0.250; Qr / Qs rows row-major instead of blocked: -0.014 ms here, +0.033 ms in each of
// node_encoder / node_update.
#ifndef AGX_A16_NODES
#define AGX_A16_NODES 5     // receivers per CTA group
#define AGX_A16_CTAS 4      // resident CTAs per SM
#endif
constexpr int A16_NODES = AGX_A16_NODES;
constexpr int A16_LANES = BLK_COLS / 4;                // 38
constexpr int A16_THREADS = A16_NODES * A16_LANES;     // 304
// Relations per receiver per task.  All A16_BATCH slots are always evaluated (branch-free), so a slot past the receiver's degree
// still costs its arithmetic and a (stale) sender-row gather from L2: cloth's receivers have exactly topk + tools = 7 relations,
// and with 7 slots instead of 8 nothing is wasted on them (0.213 -> 0.205 ms; granular-1k x 64: 0.098 -> 0.094; r02Y).  With 7
// slots the kernel fits 71 registers and 20 receivers per SM pay -- 5 CTAs x 4 receivers 0.194 ms, 4 CTAs x 5 receivers (shipped)
// the same on cloth and 2 % more rollout throughput on granular over the whole configs[4] sweep (r02K); at 8 slots 4 x 4 was best.
#ifndef AGX_A16_BATCH
#define AGX_A16_BATCH 7
#endif
constexpr int A16_BATCH = AGX_A16_BATCH;
constexpr int A16_CBUF = A16_NODES * A16_BATCH * C16_ROW;   // bytes per stage
constexpr size_t A16_SMEM = 2 * (size_t)A16_CBUF + 128;

__device__ __forceinline__ uint2 lds64u(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ int lds_s8(uint32_t addr) {
  int v;
  asm volatile("ld.shared.s8 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ int lds_s32(uint32_t addr) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

__global__ void __launch_bounds__(A16_THREADS, AGX_A16_CTAS) edge_aggregate_c16_kernel(
    const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ send, int rows, int N, uint32_t N_magic, int E_cap,
    const uint8_t* __restrict__ C16, const float4* __restrict__ Qr, const float4* __restrict__ Qs, float4* __restrict__ agg) {
  extern __shared__ uint8_t a16_raw[];
  constexpr int IDS = (A16_BATCH + 3) / 4 * 4;   // padded: a slot's ids are read with 128-bit loads
  __shared__ __align__(16) int32_t ids[2][A16_NODES][IDS];
  __shared__ __align__(8) uint64_t bar[2];
  uint8_t* cbuf = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(a16_raw) + 127) & ~(uintptr_t)127);   // [2][NODES][BATCH][C16_ROW]
  const int slot = threadIdx.x / A16_LANES, j = threadIdx.x - slot * A16_LANES;
  for (int i = threadIdx.x; i < 2 * A16_NODES * IDS; i += A16_THREADS) (&ids[0][0][0])[i] = 0;
  for (int i = threadIdx.x; i < 2 * A16_CBUF / 16; i += A16_THREADS) reinterpret_cast<uint4*>(cbuf)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], A16_NODES);
    mbar_init(&bar[1], A16_NODES);
    fence_mbar_init();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the zero fill (generic proxy) is ordered before the bulk copies (async proxy)
  __syncthreads();
  pdl_wait();      // the shared-memory set-up above overlaps the previous kernel's tail
  pdl_trigger();
  const int n_groups = (rows + A16_NODES - 1) / A16_NODES, G = gridDim.x;
  // this thread's 4 columns [4j, 4j + 4): a quarter of piece p = j >> 2 (the narrow last piece has two quarters)
  const int p = j >> 2;
  const bool narrow = 4 * j >= BLK_LAST;
  // shared-memory byte addresses of this thread's data inside stage 0 (stage 1: + A16_CBUF / + sizeof(ids[0]))
  const uint32_t c_at = smem_u32(cbuf) + (uint32_t)slot * (A16_BATCH * C16_ROW) + 8u * j;
  const uint32_t e_at = smem_u32(cbuf) + (uint32_t)slot * (A16_BATCH * C16_ROW) + C16_EXP_OFF + 8 * (p & 1) + (p >> 1);
  const uint32_t id_at = smem_u32(&ids[0][slot][0]);
  // float4 index of the thread's columns in the blocked fp32 rows (tc_chain.cuh: blk_off): row * jrow + (row / 128) * jtile + joff
  const uint32_t jrow = narrow ? (BLK_COLS - BLK_LAST) / 4 : BLK_W / 4;
  const uint32_t jtile = BLK_TILE / 4 - TILE * jrow;
  const uint32_t joff = (uint32_t)p * (TILE * BLK_W / 4) + (j & 3);
  auto idx = [=](uint32_t row) { return row * jrow + (row >> 7) * jtile + joff; };
  auto at = [=](const float4* m, uint32_t row) { return __ldg(m + idx(row)); };
  // row bounds of this thread's receiver in group v, as loaded (clamped to the capacity where they are consumed: the loads are
  // issued two groups ahead and nothing may wait for them before the end of the iteration)
  auto load_bounds = [&](int v, int& beg, int& end) {
    const int r = v * A16_NODES + slot;
    beg = end = 0;
    if (v < n_groups && r < rows) { beg = __ldg(row_ptr + r); end = __ldg(row_ptr + r + 1); }
  };
  // relations [b0, e0) of this slot (at most BATCH): sender ids by cp.async (threads j < BATCH), their C16 rows by ONE bulk copy
  auto stage = [&](int buf, int b0, int e0) {
    if (j < A16_BATCH && b0 + j < e0)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(&ids[buf][slot][j])), "l"(send + b0 + j) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (j == 0) {
      const int n = min(e0 - b0, A16_BATCH);
      if (n > 0) {
        mbar_arrive_expect_tx(&bar[buf], (uint32_t)n * C16_ROW);
        bulk_g2s(cbuf + (size_t)buf * A16_CBUF + (size_t)slot * A16_BATCH * C16_ROW, C16 + (size_t)b0 * C16_ROW, (uint32_t)n * C16_ROW, &bar[buf]);
      } else {
        mbar_arrive(&bar[buf]);
      }
    }
  };
  // acc += m * relu((c + qr) + qs),  c = (u16 - 32768) * 2^-e as one FMA on the float whose mantissa holds u16 (packed fp32 pairs)
  auto accumulate = [&](const uint2 w, int e, const float4 qr, const float4 qs, float m, float4& acc) {
    const float sc = __uint_as_float((uint32_t)(127 - e) << 23), off = -C16_MAGIC * sc;
    const float2 sc2 = make_float2(sc, sc), off2 = make_float2(off, off), m2 = make_float2(m, m);
    float2 c0 = __ffma2_rn(make_float2(__uint_as_float(__byte_perm(w.x, 0x4B000000u, 0x7410)), __uint_as_float(__byte_perm(w.x, 0x4B000000u, 0x7432))), sc2, off2);
    float2 c1 = __ffma2_rn(make_float2(__uint_as_float(__byte_perm(w.y, 0x4B000000u, 0x7410)), __uint_as_float(__byte_perm(w.y, 0x4B000000u, 0x7432))), sc2, off2);
    c0 = __fadd2_rn(__fadd2_rn(c0, make_float2(qr.x, qr.y)), make_float2(qs.x, qs.y));
    c1 = __fadd2_rn(__fadd2_rn(c1, make_float2(qr.z, qr.w)), make_float2(qs.z, qs.w));
    const float2 a0 = __ffma2_rn(make_float2(fmaxf(c0.x, 0.f), fmaxf(c0.y, 0.f)), m2, make_float2(acc.x, acc.y));
    const float2 a1 = __ffma2_rn(make_float2(fmaxf(c1.x, 0.f), fmaxf(c1.y, 0.f)), m2, make_float2(acc.z, acc.w));
    acc = make_float4(a0.x, a0.y, a1.x, a1.y);
  };

  // A task is (receiver group v, batch k): relations [beg + k * BATCH, +BATCH) of every receiver of the group.  Receivers with more
  // than BATCH relations (granular: up to topk + tools = 25) take further batches through the same pipeline; `more` (CTA-uniform,
  // from the barrier that ends the previous iteration) says whether the current group has relations left after the current batch.
  int v = blockIdx.x, k = 0;
  int beg, end, beg1, end1, raw2b = 0, raw2e = 0;
  uint32_t parity = 0;   // bit b: phase to wait for on bar[b]
  load_bounds(v, beg, end);
  load_bounds(v + G, beg1, end1);
  beg = min(beg, E_cap); end = min(end, E_cap); beg1 = min(beg1, E_cap); end1 = min(end1, E_cap);
  stage(0, beg, end);
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  int more = __syncthreads_or(end - beg > A16_BATCH);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 qr = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int it = 0; v < n_groups; ++it) {
    const int buf = it & 1;
    const int r = v * A16_NODES + slot;
    const bool valid = r < rows;
    float4 q[A16_BATCH];
    const int n0 = min(end - (beg + k * A16_BATCH), A16_BATCH);          // relations of this batch (<= 0: none)
    uint32_t gb = 0;
    if (valid) {   // first particle of the receiver's graph: r - r % N with a multiply-high (one correction step either way)
      uint32_t qd = __umulhi((uint32_t)r, N_magic);
      int rem = r - (int)(qd * (uint32_t)N);
      rem += rem < 0 ? N : 0;
      rem -= rem >= N ? N : 0;
      gb = (uint32_t)(r - rem);
    }
    if (k == 0 && valid) qr = at(Qr, (uint32_t)r);
    {
      const uint32_t ida = id_at + buf * (uint32_t)sizeof(ids[0]);
#pragma unroll
      for (int u = 0; u < A16_BATCH; ++u) q[u] = at(Qs, gb + (uint32_t)lds_s32(ida + 4 * u));   // slots past the degree: a stale (valid) id
    }
    // the next task's index data and C rows; row bounds of the group after the next (consumed at the end of the iteration)
    const int nk = more ? k + 1 : 0;
    const int nbeg = more ? beg : beg1, nend = more ? end : end1;
    stage(buf ^ 1, nbeg + nk * A16_BATCH, nend);
    if (!more) load_bounds(v + 2 * G, raw2b, raw2e);
    mbar_wait(&bar[buf], (parity >> buf) & 1);     // this batch's C16 rows have landed
    parity ^= 1u << buf;
    {
      const uint32_t ca = c_at + buf * A16_CBUF, ea = e_at + buf * A16_CBUF;
#pragma unroll
      for (int u = 0; u < A16_BATCH; ++u)
        accumulate(lds64u(ca + u * C16_ROW), lds_s8(ea + u * C16_ROW), qr, q[u], u < n0 ? 1.f : 0.f, acc);
    }
    // the finished row leaves as plain fp32 (blocked like Qr / Qs): the update chain takes the row maximum, scales and splits it for
    // its tensor-memory A (tc_node_update_kernel<.., AGG32>).  This kernel is bound by instruction issue and by the length of a
    // task's dependent chain: the row maximum (match / redux / shared-memory atomic, read back after the barrier), the power-of-two
    // scale and the fp16 hi / lo conversion were a fifth of its instructions -- 0.249 -> 0.213 ms on cloth-2k x 128 (r02U), 4.9 TB/s.
    // Leaving just two partial maxima per row behind (one REDUX per warp segment, 12 instructions) already costs 0.016 ms (r02V).
    if (!more && valid) agg[idx((uint32_t)r)] = acc;
    asm volatile("cp.async.wait_group 0;" ::: "memory");   // the next task's sender ids have landed ...
    // ... and are visible; everyone is done with cbuf[buf]; does the next task's group go on after it?
    const int more_next = __syncthreads_or(nend - nbeg > (nk + 1) * A16_BATCH);
    if (!more) {
      acc = make_float4(0.f, 0.f, 0.f, 0.f);
      asm volatile("" : "+r"(raw2b), "+r"(raw2e));         // the row_ptr loads issued above are first waited for HERE
      beg = beg1; end = end1; beg1 = min(raw2b, E_cap); end1 = min(raw2e, E_cap);
      v += G;
    }
    k = nk;
    more = more_next;
  }
}

}  // namespace tc

// ------------------------------------------------------------------------------------ host entry points (used by forward.cu)
size_t tc_blob_bytes(size_t base_bytes) { return tc::tc_layout(base_bytes).total; }

int tc_pack(const AgxModelDims* dims, const AgxWeights* raw, void* packed, size_t base_bytes, cudaStream_t st) {
  using namespace tc;
  const int F = dims->F;
  const int d_node = dims->d_attr + dims->d_phys + dims->d_act, d_rel = 2 * dims->d_attr + 1 + 3 * dims->n_his;
  // the kernels place the bias constants at fixed A columns (6 / 17 / ONE_COL)
  AGX_REQUIRE(d_node == 6 && d_rel == 17 && F == ONE_COL, AGX_ERR_ARG, "tc_pack: the tensor-core path is built for 6/17/150 input dims");
  PackArgs a;
  auto set = [&](int t, int layer, int ld, int col0, int K, bool bias) {
    a.spec[t] = PackSpec{raw->weight[layer], bias ? raw->bias[layer] : nullptr, ld, col0, K, F, 0};
  };
  // transposed: n_out = columns of the original block (the layer's inputs), K = its F rows
  auto set_t = [&](int t, int layer, int ld, int col0, int n_out) {
    a.spec[t] = PackSpec{raw->weight[layer], nullptr, ld, col0, F, n_out, 1};
  };
  set(T_PENC0, AGX_W_PENC0, d_node, 0, d_node, true);
  set(T_PENC2, AGX_W_PENC2, F, 0, F, true);
  set(T_PENC4, AGX_W_PENC4, F, 0, F, true);
  set(T_RENC0, AGX_W_RENC0, d_rel, 0, d_rel, true);
  set(T_RENC2, AGX_W_RENC2, F, 0, F, true);
  set(T_RENC4, AGX_W_RENC4, F, 0, F, true);
  set(T_RP_REL, AGX_W_RPROP, 3 * F, 0, F, true);
  set(T_RP_RECV, AGX_W_RPROP, 3 * F, F, F, false);
  set(T_RP_SEND, AGX_W_RPROP, 3 * F, 2 * F, F, false);
  set(T_PP_ENC, AGX_W_PPROP, 2 * F, 0, F, true);
  set(T_PP_AGG, AGX_W_PPROP, 2 * F, F, F, false);
  set(T_PRED0, AGX_W_PRED0, F, 0, F, true);
  set(T_PRED1, AGX_W_PRED1, F, 0, F, true);
  set_t(TT_PENC2, AGX_W_PENC2, F, 0, F);
  set_t(TT_PENC4, AGX_W_PENC4, F, 0, F);
  set_t(TT_RENC0, AGX_W_RENC0, d_rel, 0, d_rel);
  set_t(TT_RENC2, AGX_W_RENC2, F, 0, F);
  set_t(TT_RENC4, AGX_W_RENC4, F, 0, F);
  set_t(TT_RP_REL, AGX_W_RPROP, 3 * F, 0, F);
  set_t(TT_RP_RECV, AGX_W_RPROP, 3 * F, F, F);
  set_t(TT_RP_SEND, AGX_W_RPROP, 3 * F, 2 * F, F);
  set_t(TT_PP_ENC, AGX_W_PPROP, 2 * F, 0, F);
  set_t(TT_PP_AGG, AGX_W_PPROP, 2 * F, F, F);
  set_t(TT_PRED0, AGX_W_PRED0, F, 0, F);
  set_t(TT_PRED1, AGX_W_PRED1, F, 0, F);
  a.blob = static_cast<uint8_t*>(packed);
  a.L = tc_layout(base_bytes);
  { ProfScope ps(AGX_KIND_OTHER, st);
    pack_tc_kernel<<<T_NUM, 256, 0, st>>>(a); }
  AGX_LAUNCH_CHECK();
  return AGX_OK;
}

static int tc_ensure_attrs() {
  static thread_local DeviceOnce once;
  if (!once.need()) return AGX_OK;
  using namespace tc;
#ifdef AGX_TC_STAGGER
  if (const char* e = getenv("AGX_TC_STAGGER")) {
    const int v = atoi(e);
    AGX_CUDA_OK(cudaMemcpyToSymbol(g_tc_stagger, &v, sizeof(v)));
  }
#endif
  AGX_CUDA_OK(cudaFuncSetAttribute(tc_edge_encoder_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  AGX_CUDA_OK(cudaFuncSetAttribute(tc_edge_encoder_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  AGX_CUDA_OK(cudaFuncSetAttribute(edge_aggregate_c16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)A16_SMEM));
  AGX_CUDA_OK(cudaFuncSetAttribute(tc_node_encoder_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  AGX_CUDA_OK(cudaFuncSetAttribute(tc_node_update_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  AGX_CUDA_OK(cudaFuncSetAttribute(tc_node_update_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  AGX_CUDA_OK(cudaFuncSetAttribute((tc_edge_encoder_kernel<false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  AGX_CUDA_OK(cudaFuncSetAttribute(tc_node_encoder_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  AGX_CUDA_OK(cudaFuncSetAttribute((tc_node_update_kernel<false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  AGX_CUDA_OK(cudaFuncSetAttribute((tc_node_update_kernel<true, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  AGX_CUDA_OK(cudaFuncSetAttribute(tc_lin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  AGX_CUDA_OK(cudaFuncSetAttribute((tc_node_update_kernel<false, false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  AGX_CUDA_OK(cudaFuncSetAttribute((tc_node_update_kernel<true, false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  return AGX_OK;
}

int tc_node_encoder(const AgxGraphIn* g, const float* wts, const PackedLayout& PL, size_t base_bytes, const TcFwdBuffers& w, cudaStream_t st) {
  using namespace tc;
  if (int rc = tc_ensure_attrs()) return rc;
  const int64_t rows = (int64_t)g->B * g->N;
  const int tiles = (int)((rows + TILE - 1) / TILE);
  NodeArgs a{g->state, g->attrs, g->action, g->p_instance, g->physics, g->B, g->N, g->n_p,
             reinterpret_cast<const uint8_t*>(wts), tc_layout(base_bytes),
             w.nfeat, w.P0, w.A, w.Qr0, w.Qs0, w.rowmaxP0, w.rowmaxA, w.S0, nullptr, nullptr, nullptr, nullptr};
  if (w.save) { a.sv_p_in = w.save->p_in; a.sv_h1 = w.save->h1; a.sv_h2 = w.save->h2; a.sv_penc = w.save->penc; }
  { ProfScope ps(AGX_KIND_NODE_ENCODER, st);
    if (w.save) launch_pdl(PDL_BIG, tc_node_encoder_kernel<true>, tiles < num_sms() ? tiles : num_sms(), THREADS, SMEM_BYTES, st, a);
    else launch_pdl(PDL_BIG, tc_node_encoder_kernel<false>, tiles < num_sms() ? tiles : num_sms(), THREADS, SMEM_BYTES, st, a); }
  AGX_LAUNCH_CHECK();
  return AGX_OK;
}

// history records only (later steps of a rollout: the encoder's products are reused)
int tc_nfeat(const AgxGraphIn* g, const TcFwdBuffers& w, cudaStream_t st) {
  using namespace tc;
  const int64_t rows = (int64_t)g->B * g->N;
  NodeArgs a{g->state, g->attrs, g->action, g->p_instance, g->physics, g->B, g->N, g->n_p, nullptr, TcLayout{},
             w.nfeat, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  { ProfScope ps(AGX_KIND_NODE_ENCODER, st);
    nfeat_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(a); }
  AGX_LAUNCH_CHECK();
  return AGX_OK;
}

int tc_edge_encoder(const AgxGraphIn* g, const float* wts, const PackedLayout& PL, size_t base_bytes, const TcFwdBuffers& w, bool mixed,
                    cudaStream_t st) {
  using namespace tc;
  if (int rc = tc_ensure_attrs()) return rc;
  const int64_t rows = (int64_t)g->B * g->N;
  const int64_t tiles = (g->E_cap + TILE - 1) / TILE;
  EdgeArgs a{g->row_ptr, g->send, g->recv, rows, g->N, g->E_cap, w.nfeat, reinterpret_cast<const uint8_t*>(wts), tc_layout(base_bytes),
             w.C, nullptr, nullptr, nullptr, nullptr};
  AGX_REQUIRE(!(w.save && mixed), AGX_ERR_ARG, "edge_encoder: the training copies need the fp32 relation term (AGX_PREC_TC_F16X3)");
  if (w.save) { a.sv_rel_in = w.save->rel_in; a.sv_g1 = w.save->g1; a.sv_g2 = w.save->g2; a.sv_renc = w.save->renc; }
  { ProfScope ps(AGX_KIND_EDGE_ENCODER, st);
    const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
    if (w.save) launch_pdl(PDL_BIG, tc_edge_encoder_kernel<false, true>, grid, THREADS, SMEM_BYTES, st, a);
    else if (mixed) launch_pdl(PDL_BIG, tc_edge_encoder_kernel<true>, grid, THREADS, SMEM_BYTES, st, a);
    else launch_pdl(PDL_BIG, tc_edge_encoder_kernel<false>, grid, THREADS, SMEM_BYTES, st, a); }
  AGX_LAUNCH_CHECK();
  return AGX_OK;
}

// Tensor-core replacement of train.cu's lin<160> (see tc_lin_kernel); `packed` is the blob written by agx_pack_weights.
int tc_lin(cudaStream_t st, const void* packed, size_t base_bytes, int layer, const float* X, int ldx, const float* mask, int ldm,
           const float* bias, const float* add1, const float* add2, float* Y, int ldy, int64_t M, bool relu, bool accumulate, int n_store) {
  using namespace tc;
  if (M <= 0) return AGX_OK;
  if (int rc = tc_ensure_attrs()) return rc;
  AGX_REQUIRE(layer >= 0 && layer < T_NUM && tc_kpad(layer) == FP && n_store % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && (!mask || ldm % 4 == 0),
              AGX_ERR_ARG, "tc_lin: unsupported layer / strides");
  LinTcArgs a{X, ldx, mask, ldm, bias, add1, add2, Y, ldy, M, relu ? 1 : 0, accumulate ? 1 : 0, n_store, layer,
              reinterpret_cast<const uint8_t*>(packed), tc_layout(base_bytes)};
  const int64_t tiles = (M + TILE - 1) / TILE;
  { ProfScope ps(AGX_KIND_OTHER, st);
    tc_lin_kernel<<<(int)(tiles < num_sms() ? tiles : num_sms()), THREADS, SMEM_BYTES, st>>>(a); }
  AGX_LAUNCH_CHECK();
  return AGX_OK;
}

int tc_edge_aggregate(const AgxGraphIn* g, const TcFwdBuffers& w, bool mixed, bool first, cudaStream_t st) {
  using namespace tc;
  const int64_t rows = (int64_t)g->B * g->N;
  const float* Qr = first ? w.Qr0 : w.Qr;   // propagation step 0 gathers the encoder's products
  const float* Qs = first ? w.Qs0 : w.Qs;
  if (mixed) {
    if (int rc = tc_ensure_attrs()) return rc;
    AGX_REQUIRE(w.agg_f32, AGX_ERR_ARG, "edge_aggregate: the C16 aggregate writes plain fp32 rows (TcFwdBuffers::agg_f32)");
    // the kernel indexes the blocked fp32 rows with 32-bit float offsets
    AGX_REQUIRE(blk_rows(rows) * (BLK_COLS / 4) < (1ll << 32) && g->E_cap < (1ll << 31), AGX_ERR_ARG,
                "edge_aggregate: %lld rows / %lld relations exceed the 32-bit feature index", (long long)rows, (long long)g->E_cap);
    const int64_t groups = (rows + A16_NODES - 1) / A16_NODES, resident = (int64_t)num_sms() * AGX_A16_CTAS;
    { ProfScope ps(AGX_KIND_EDGE_AGGREGATE, st);
      launch_pdl(PDL_BIG, edge_aggregate_c16_kernel, (unsigned)(groups < resident ? groups : resident), A16_THREADS, A16_SMEM, st, 
          g->row_ptr, g->send, (int)rows, g->N, g->N == 1 ? 0xffffffffu : (uint32_t)((1ull << 32) / (uint64_t)g->N), (int)g->E_cap,
          reinterpret_cast<const uint8_t*>(w.C),
          reinterpret_cast<const float4*>(Qr),
          reinterpret_cast<const float4*>(Qs), reinterpret_cast<float4*>(w.agg)); }
    AGX_LAUNCH_CHECK();
    return AGX_OK;
  }
  // the kernel indexes float4s with 32 bits: 40 per row (68 GB of features per buffer at the limit)
  AGX_REQUIRE(blk_rows(rows) * (BLK_COLS / 4) < (1ll << 32) && blk_rows(g->E_cap) * (BLK_COLS / 4) < (1ll << 32) && g->E_cap < (1ll << 31), AGX_ERR_ARG,
              "edge_aggregate: %lld rows / %lld relations exceed the 32-bit feature index", (long long)rows, (long long)g->E_cap);
  const int64_t groups = (rows + AGG_NODES - 1) / AGG_NODES, resident = (int64_t)num_sms() * AGG_CTAS_PER_SM;
  { ProfScope ps(AGX_KIND_EDGE_AGGREGATE, st);
    launch_pdl(PDL_BIG, edge_aggregate_split_kernel, (unsigned)(groups < resident ? groups : resident), AGG_THREADS, 0, st, 
        g->row_ptr, g->send, (int)rows, g->N, (int)g->E_cap, reinterpret_cast<const float4*>(w.C), reinterpret_cast<const float4*>(Qr),
        reinterpret_cast<const float4*>(Qs), reinterpret_cast<uint32_t*>(w.agg), w.agg_exp, w.agg_max,
        w.save ? reinterpret_cast<float4*>(w.save->agg_f32) : nullptr); }
  AGX_LAUNCH_CHECK();
  return AGX_OK;
}

int tc_node_update(const AgxGraphIn* g, const float* wts, const PackedLayout& PL, size_t base_bytes, const TcFwdBuffers& w, bool first,
                   bool last, float* pred_pos, int64_t pos_stride_b, float* pred_motion, cudaStream_t st) {
  using namespace tc;
  if (int rc = tc_ensure_attrs()) return rc;
  const int64_t rows = (int64_t)g->B * g->N;
  const int tiles = (int)((rows + TILE - 1) / TILE);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  UpdArgs a{g->B, g->N, g->n_p, reinterpret_cast<const uint32_t*>(w.agg), w.agg_exp, w.agg_max, w.A, first ? w.P0 : w.P, w.P, w.Qr, w.Qs,
            first ? w.rowmaxP0 : w.rowmaxP, w.rowmaxP, w.rowmaxA, first ? w.S0 : nullptr,
            reinterpret_cast<const uint8_t*>(wts), tc_layout(base_bytes),
            wts + PL.pred2_w, g->state, pred_pos, pos_stride_b, pred_motion, nullptr, nullptr, nullptr};
  if (w.save) { a.sv_P = w.save->P_next; a.sv_u1 = w.save->u1; a.sv_u2 = w.save->u2; }
  if (last) {
    ProfScope ps(AGX_KIND_NODE_HEAD, st);
    if (w.save) launch_pdl(PDL_BIG, tc_node_update_kernel<true, true>, grid, THREADS, SMEM_BYTES, st, a);
    else if (w.agg_f32) launch_pdl(PDL_BIG, (tc_node_update_kernel<true, false, true>), grid, THREADS, SMEM_BYTES, st, a);
    else launch_pdl(PDL_BIG, tc_node_update_kernel<true>, grid, THREADS, SMEM_BYTES, st, a);
  } else {
    ProfScope ps(AGX_KIND_NODE_UPDATE, st);
    if (w.save) launch_pdl(PDL_BIG, tc_node_update_kernel<false, true>, grid, THREADS, SMEM_BYTES, st, a);
    else if (w.agg_f32) launch_pdl(PDL_BIG, (tc_node_update_kernel<false, false, true>), grid, THREADS, SMEM_BYTES, st, a);
    else launch_pdl(PDL_BIG, tc_node_update_kernel<false>, grid, THREADS, SMEM_BYTES, st, a);
  }
  AGX_LAUNCH_CHECK();
  return AGX_OK;
}

}  // namespace agx

#include "tc_backward.inl"

#ifdef AGX_TC_TIMELINE
// Debug-only entry (tools/tc_timeline.py builds a separate library with -DAGX_TC_TIMELINE): points CTA 0's MMA thread at a
// device buffer of 4096 int64 stamps (slot << 48 | clock64).
extern "C" __attribute__((visibility("default"))) int agx_debug_set_timeline(void* dev_ptr) {
  long long* p = static_cast<long long*>(dev_ptr);
  return cudaMemcpyToSymbol(agx::tc::g_tc_timeline, &p, sizeof(p)) == cudaSuccess ? 0 : -3;
}
#endif
