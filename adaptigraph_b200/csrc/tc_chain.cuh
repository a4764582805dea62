// Tensor-core MLP chains (AGX_PREC_TC_F16X3): tcgen05.mma kind::f16 with fp32-accurate split operands.
//
// One CTA owns a tile of 128 rows (relations or particles) and pushes it through a chain of dense
// layers without the activations ever leaving the SM:
//
//   * accumulators live in tensor memory (two 160-column fp32 buffers, ping-pong across layers);
//   * the layer input A also lives in tensor memory (TS-mode MMA): every fp32 activation a is scaled
//     by an exact per-row power of two and split into two fp16 values a' = hi + lo (22 significant
//     bits), stored as packed half2 columns (80 columns for hi, 80 for lo);
//   * the weights are pre-packed (agx_pack_weights) as scaled fp16 hi/lo images in the canonical
//     K-major core-matrix layout and brought into shared memory by the TMA engine as one bulk copy
//     per layer, double buffered so the next layer's weights land while the current layer runs;
//   * a layer is 3 MMAs per 16-wide K step:  D += Alo*Whi + Ahi*Wlo + Ahi*Whi  (the dropped lo*lo
//     term is 2^-22 relative), issued by one thread; fp32 accumulation in tensor memory;
//   * 16 epilogue warps (thread = row; the four warps of a 32-lane quarter split every 32-column
//     chunk into 8-column pieces) read the accumulator with tcgen05.ld (whole row piece up front, one
//     wait, buffer released immediately), undo the power-of-two scales exactly, apply bias /
//     residual / ReLU, and either re-split the result into the next layer's A (tcgen05.st) chunk by
//     chunk — the MMA warp starts the next layer's K steps as soon as a chunk is complete — or store
//     fp32 rows to HBM with 256-bit stores.
//
// Warp roles: warps 0-15 epilogue + input producers, warp 16 MMA issuer (+ TMEM allocation),
// warp 17 weight loader.  All waits are bounded (tc_ptx.cuh: a protocol bug traps, it cannot hang).
#pragma once
#include "common.cuh"
#include "tc_ptx.cuh"

namespace agx {
namespace tc {

constexpr int TILE = 128;
constexpr int NQ = 4;                                   // column quarters per chunk (warps per lane quarter)
constexpr int QW = 8;                                   // columns per thread per chunk
constexpr int EPI_WARPS = 4 * NQ;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int MMA_WARP = EPI_WARPS;
constexpr int LOAD_WARP = EPI_WARPS + 1;
constexpr int THREADS = EPI_THREADS + 64;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t COL_ACC0 = 0, COL_ACC1 = 160, COL_AHI = 320, COL_ALO = 400;
constexpr int NCHUNK = 5;                               // 32-wide K chunks of a 160-wide layer
constexpr uint32_t IMG_BIG = FP * FP * 2;               // bytes of one fp16 image (hi or lo) of a 160x160 layer
constexpr uint32_t IDESC = make_idesc_f16(TILE, FP);
constexpr int TARGET_EXP = 14;                          // scaled operands satisfy |x| <= 2^14 (fp16 max is 65504)

// tensor-core layer ids (order of the fp16 images in the packed blob)
enum TcLayer {
  T_PENC0 = 0, T_PENC2, T_PENC4, T_RENC0, T_RENC2, T_RENC4, T_RP_REL, T_RP_RECV, T_RP_SEND, T_PP_ENC, T_PP_AGG,
  T_PRED0, T_PRED1, T_NUM
};
__host__ __device__ constexpr int tc_kpad(int t) { return t == T_PENC0 ? 16 : (t == T_RENC0 ? 32 : FP); }

struct TcLayout {
  size_t meta;          // byte offset of float4 meta[T_NUM] = {2^-sw, inf-norm of W, max|bias|, 0}
  size_t img[T_NUM];    // byte offset of the hi image; the lo image follows at + FP*kpad*2
  size_t total;         // bytes
};
inline TcLayout tc_layout(size_t base_bytes) {
  TcLayout L;
  size_t o = align_up(base_bytes, 256);
  L.meta = o;
  o += T_NUM * 16;
  for (int t = 0; t < T_NUM; ++t) {
    o = align_up(o, 256);
    L.img[t] = o;
    o += 2 * (size_t)FP * tc_kpad(t) * 2;
  }
  L.total = align_up(o, 256);
  return L;
}

// byte offset of element (n, k) inside a K-major no-swizzle core-matrix image with kpad columns
__host__ __device__ constexpr uint32_t img_offset(int n, int k, int kpad) {
  return (uint32_t)((n >> 3) * (kpad >> 3) * 128 + (k >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2);
}

// exponent e such that bound * 2^e <= 2^TARGET_EXP (exact power-of-two scaling)
__device__ __forceinline__ int scale_exp(float bound) {
  if (!(bound > 0.f)) return 0;
  const int x = (int)((__float_as_uint(bound) >> 23) & 0xff) - 126;   // bound <= 2^x
  return max(-100, min(100, TARGET_EXP - x));
}
__device__ __forceinline__ float exp2i(int e) { return __uint_as_float((uint32_t)(e + 127) << 23); }

// split 8 scaled fp32 values into packed fp16 hi / lo columns (hi = round-to-nearest, lo = residual)
__device__ __forceinline__ void split8(const float (&s)[QW], uint32_t (&hi)[4], uint32_t (&lo)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 h = __floats2half2_rn(s[2 * i], s[2 * i + 1]);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(s[2 * i] - hf.x, s[2 * i + 1] - hf.y);
    hi[i] = *reinterpret_cast<const uint32_t*>(&h);
    lo[i] = *reinterpret_cast<const uint32_t*>(&l);
  }
}

// ------------------------------------------------------------------------------------------------
struct Shared {
  uint8_t* wbig[2];      // two 2*IMG_BIG weight buffers (hi image then lo image)
  uint8_t* wsmall;       // first-layer weights (K = 16 or 32)
  float* bias;           // [MAX_BIAS][FP]
  float* xchg;           // [NQ][TILE] row exchange between the column quarters
  float* head_w;         // [3][FP] + [4] (head program only)
  uint64_t* bar_wsmall;
  uint64_t* bar_wfull;   // [2]
  uint64_t* bar_wempty;  // [2]
  uint64_t* bar_a;       // [NCHUNK]
  uint64_t* bar_accfull; // [2]
  uint64_t* bar_accempty;// [2]
  uint32_t* tmem_ptr;
};
constexpr int MAX_BIAS = 4;
constexpr size_t SMEM_BYTES = 2 * (2 * (size_t)IMG_BIG) + 2 * (size_t)FP * 32 * 2 + MAX_BIAS * FP * 4 + NQ * TILE * 4 + (3 * FP + 4) * 4 +
                              16 * 8 + 16 + 128 /* alignment slack */;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB shared-memory limit of a CTA");

__device__ __forceinline__ Shared carve_shared(uint8_t* raw) {
  uint8_t* p = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 127) & ~(uintptr_t)127);
  Shared s;
  s.wbig[0] = p; p += 2 * IMG_BIG;
  s.wbig[1] = p; p += 2 * IMG_BIG;
  s.wsmall = p; p += 2 * FP * 32 * 2;
  s.bias = reinterpret_cast<float*>(p); p += MAX_BIAS * FP * 4;
  s.xchg = reinterpret_cast<float*>(p); p += NQ * TILE * 4;
  s.head_w = reinterpret_cast<float*>(p); p += (3 * FP + 4) * 4;
  uint64_t* b = reinterpret_cast<uint64_t*>(p);
  s.bar_wsmall = b; s.bar_wfull = b + 1; s.bar_wempty = b + 3; s.bar_a = b + 5; s.bar_accfull = b + 10; s.bar_accempty = b + 12;
  s.tmem_ptr = reinterpret_cast<uint32_t*>(b + 14);
  return s;
}

// Program description (compile time): which layers, how many K steps, whether the layer's A comes
// from the previous epilogue / producer (wait on bar_a) or is reused from the previous layer.
struct LayerStep { int layer; int ksteps; int waits_a; };

// ------------------------------------------------------------------------------------------------ roles
// Weight loader: one thread streams the big layers of every tile through the two-buffer ring.
template <int NL>
__device__ __forceinline__ void loader_role(const Shared& sh, const LayerStep (&prog)[NL], const uint8_t* blob, const TcLayout& L,
                                            int n_tiles) {
  if (!elect_one()) return;
  if ((int)blockIdx.x >= n_tiles) return;
  {
    const int t = prog[0].layer;
    if (prog[0].ksteps < 10) {
      const uint32_t bytes = 2u * FP * tc_kpad(t) * 2u;
      mbar_arrive_expect_tx(sh.bar_wsmall, bytes);
      bulk_g2s(sh.wsmall, blob + L.img[t], bytes, sh.bar_wsmall);
    }
  }
  uint32_t buf = 0, empty_parity = 0x3;   // bit b = parity to wait for on bar_wempty[b] (starts at 1: free)
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      if (prog[l].ksteps < 10) continue;
      mbar_wait(&sh.bar_wempty[buf], (empty_parity >> buf) & 1);
      empty_parity ^= 1u << buf;
      mbar_arrive_expect_tx(&sh.bar_wfull[buf], 2 * IMG_BIG);
      bulk_g2s(sh.wbig[buf], blob + L.img[prog[l].layer], 2 * IMG_BIG, &sh.bar_wfull[buf]);
      buf ^= 1;
    }
  }
}

// MMA issuer: one thread, per layer 3 MMAs per K step, commits to the accumulator / weight barriers.
template <int NL>
__device__ __forceinline__ void mma_role(const Shared& sh, const LayerStep (&prog)[NL], uint32_t tmem_base, int n_tiles) {
  if (!elect_one()) return;
  uint32_t buf = 0, full_parity = 0, accempty_parity = 0x3, a_parity = 0;
  bool small_ready = false;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      const int ksteps = prog[l].ksteps;
      const bool big = ksteps == 10;
      const int kpad = ksteps * 16;
      uint32_t w_addr;
      if (big) {
        mbar_wait(&sh.bar_wfull[buf], (full_parity >> buf) & 1);
        full_parity ^= 1u << buf;
        w_addr = smem_u32(sh.wbig[buf]);
      } else {
        if (!small_ready) { mbar_wait(sh.bar_wsmall, 0); small_ready = true; }
        w_addr = smem_u32(sh.wsmall);
      }
      const uint32_t img_bytes = (uint32_t)FP * kpad * 2;
      const uint32_t sbo = (uint32_t)(kpad >> 3) * 128;
      const int ab = l & 1;
      mbar_wait(&sh.bar_accempty[ab], (accempty_parity >> ab) & 1);
      accempty_parity ^= 1u << ab;
      const uint32_t d_tmem = tmem_base + (ab ? COL_ACC1 : COL_ACC0);
      for (int ks = 0; ks < ksteps; ++ks) {
        if (prog[l].waits_a && (ks & 1) == 0) {
          const int c = ks >> 1;
          mbar_wait(&sh.bar_a[c], (a_parity >> c) & 1);
          a_parity ^= 1u << c;
        }
        tc_fence_after();
        const uint32_t a_hi = tmem_base + COL_AHI + 8 * ks, a_lo = tmem_base + COL_ALO + 8 * ks;
        const uint64_t b_hi = make_b_desc(w_addr + ks * 256, 128, sbo);
        const uint64_t b_lo = make_b_desc(w_addr + img_bytes + ks * 256, 128, sbo);
        mma_f16_ts(d_tmem, a_lo, b_hi, IDESC, ks > 0);   // small terms first
        mma_f16_ts(d_tmem, a_hi, b_lo, IDESC, 1);
        mma_f16_ts(d_tmem, a_hi, b_hi, IDESC, 1);
      }
      mma_commit(&sh.bar_accfull[ab]);
      if (big) { mma_commit(&sh.bar_wempty[buf]); buf ^= 1; }
    }
  }
}

// ------------------------------------------------------------------------------------------------ epilogue helpers
struct EpiCtx {
  int row;          // row of the tile this thread owns (= TMEM lane)
  int q;            // column quarter: this thread handles columns [32c + 8q, 32c + 8q + 8) of every chunk
  int lane, warp;
  uint32_t tmem_lane_base;   // tmem_base + (lane quarter << 16)
  uint32_t accfull_parity;   // bit b: parity to wait for on bar_accfull[b]
  int e_in;         // exponent of the scale applied to the current A
  float rowmax_in;  // max |a| of the current A row (unscaled)
};

__device__ __forceinline__ void epi_signal_chunk(const Shared& sh, const EpiCtx& cx, int c) {
  tmem_wait_st();
  tc_fence_before();
  __syncwarp();
  if (cx.lane == 0) mbar_arrive(&sh.bar_a[c]);
}

// combine a per-thread partial row value across the four column quarters (max or sum)
template <bool IS_MAX>
__device__ __forceinline__ float epi_exchange(const Shared& sh, const EpiCtx& cx, float v) {
  sts32(&sh.xchg[cx.q * TILE + cx.row], v);
  named_bar_sync(1, EPI_THREADS);
  float r = IS_MAX ? v : 0.f;
#pragma unroll
  for (int j = 0; j < NQ; ++j) {
    const float o = lds32(&sh.xchg[j * TILE + cx.row]);
    r = IS_MAX ? fmaxf(r, o) : r + o;      // sums are formed in quarter order by every thread: identical result in all four
  }
  named_bar_sync(1, EPI_THREADS);
  return r;
}

// Writes this thread's 8 values of chunk c of the next layer's A.
__device__ __forceinline__ void epi_store_a(const EpiCtx& cx, int c, const float (&v)[QW], float scale) {
  float s[QW];
#pragma unroll
  for (int i = 0; i < QW; ++i) s[i] = v[i] * scale;
  uint32_t hi[4], lo[4];
  split8(s, hi, lo);
  tmem_st4(cx.tmem_lane_base + COL_AHI + 16 * c + 4 * cx.q, hi);
  tmem_st4(cx.tmem_lane_base + COL_ALO + 16 * c + 4 * cx.q, lo);
}

// Generic layer epilogue.  The thread's 5 x 8 accumulator values are read up front (one wait) and the
// accumulator buffer is handed back to the MMA warp immediately; then per chunk:
//   v = acc * unscale (+ bias) ; extra(c, col0, v) ; [relu] ; consume(c, col0, v).
template <bool RELU, class Extra, class Consume>
__device__ __forceinline__ void epi_layer(const Shared& sh, EpiCtx& cx, int ab, float unscale, const float* bias_s, Extra extra, Consume consume) {
  mbar_wait(&sh.bar_accfull[ab], (cx.accfull_parity >> ab) & 1);
  cx.accfull_parity ^= 1u << ab;
  tc_fence_after();
  const uint32_t acc = cx.tmem_lane_base + (ab ? COL_ACC1 : COL_ACC0) + QW * cx.q;
  uint32_t r[NCHUNK][QW];
#pragma unroll
  for (int c = 0; c < NCHUNK; ++c) tmem_ld8(acc + 32 * c, r[c]);
  tmem_wait_ld();
  tc_fence_before();
  __syncwarp();
  if (cx.lane == 0) mbar_arrive(&sh.bar_accempty[ab]);
#pragma unroll
  for (int c = 0; c < NCHUNK; ++c) {
    const int col0 = 32 * c + QW * cx.q;
    float v[QW];
    if (bias_s) {
      const float4 b0 = lds128(bias_s + col0), b1 = lds128(bias_s + col0 + 4);
      v[0] = fmaf(__uint_as_float(r[c][0]), unscale, b0.x); v[1] = fmaf(__uint_as_float(r[c][1]), unscale, b0.y);
      v[2] = fmaf(__uint_as_float(r[c][2]), unscale, b0.z); v[3] = fmaf(__uint_as_float(r[c][3]), unscale, b0.w);
      v[4] = fmaf(__uint_as_float(r[c][4]), unscale, b1.x); v[5] = fmaf(__uint_as_float(r[c][5]), unscale, b1.y);
      v[6] = fmaf(__uint_as_float(r[c][6]), unscale, b1.z); v[7] = fmaf(__uint_as_float(r[c][7]), unscale, b1.w);
    } else {
#pragma unroll
      for (int i = 0; i < QW; ++i) v[i] = __uint_as_float(r[c][i]) * unscale;
    }
    extra(c, col0, v);
    if (RELU) {
#pragma unroll
      for (int i = 0; i < QW; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    consume(c, col0, v);
  }
}

}  // namespace tc
}  // namespace agx
