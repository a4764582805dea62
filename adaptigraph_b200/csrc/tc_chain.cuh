// Tensor-core MLP chains (AGX_PREC_TC_F16X3): tcgen05.mma kind::f16 with fp32-accurate split operands.
//
// One CTA owns a tile of 128 rows (relations or particles) and pushes it through a chain of dense
// layers without the activations ever leaving the SM:
//
//   * the accumulator (160 fp32 columns) and the layer input A live in tensor memory (TS-mode MMA);
//     every fp32 activation a is scaled by an exact per-row power of two and split into two fp16
//     values a' = hi + lo (22 significant bits), stored as packed half2 columns (80 hi + 80 lo);
//     A is double buffered: a layer's epilogue writes the next layer's input into the other buffer;
//   * the weights are pre-packed (agx_pack_weights) as scaled fp16 hi/lo images in the canonical
//     K-major core-matrix layout and brought into shared memory by the TMA engine as one bulk copy
//     per layer, double buffered so the next layer's weights land while the current layer runs;
//   * a layer is 3 MMAs per 16-wide K step:  D += Alo*Whi + Ahi*Wlo + Ahi*Whi  (the dropped lo*lo
//     term is 2^-22 relative), issued by one thread; fp32 accumulation in tensor memory.  Each
//     layer is issued as two column parts (N = 96, then N = 64): the epilogue of part A overlaps the
//     MMAs of part B, and the next layer's part A starts on the K chunks part A's epilogue produced
//     while part B's epilogue is still running, so the tensor pipe does not drain between layers;
//   * 16 epilogue warps (thread = row; the four warps of a 32-lane quarter split every 32-column
//     chunk into 8-column pieces) read an accumulator part with tcgen05.ld (one wait, part released
//     to the MMA warp immediately), undo the power-of-two scales exactly, apply bias / residual /
//     ReLU, and either re-split the result into the next layer's A (tcgen05.st) chunk by chunk or
//     store fp32 rows to HBM with 256-bit stores;
//   * the input of the NEXT tile is fetched early and written into the free A buffer before the
//     current tile's last epilogue, so the first layer of the next tile overlaps that epilogue.
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
constexpr uint32_t COL_ACC = 0;
constexpr uint32_t COL_AHI0 = 160, COL_ALO0 = 240, COL_AHI1 = 320, COL_ALO1 = 400;
constexpr int NCHUNK = 5;                               // 32-wide K chunks of a 160-wide layer
constexpr int NCHUNK_A = 3;                             // accumulator part A = chunks 0..2 (N = 96), part B = chunks 3..4 (N = 64)
constexpr int N_PART_A = 32 * NCHUNK_A, N_PART_B = FP - N_PART_A;
constexpr uint32_t IMG_BIG = FP * FP * 2;               // bytes of one fp16 image (hi or lo) of a 160x160 layer
constexpr uint32_t IDESC_A = make_idesc_f16(TILE, N_PART_A), IDESC_B = make_idesc_f16(TILE, N_PART_B);
constexpr int TARGET_EXP = 14;                          // scaled operands satisfy |x| <= 2^14 (fp16 max is 65504)

__device__ __forceinline__ uint32_t col_ahi(int buf) { return buf ? COL_AHI1 : COL_AHI0; }
__device__ __forceinline__ uint32_t col_alo(int buf) { return buf ? COL_ALO1 : COL_ALO0; }

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

// scale 8 fp32 values and split them into packed fp16 hi / lo columns (hi = round-to-nearest, lo = residual);
// fp32 arithmetic on register pairs (FMUL2 / FADD2)
__device__ __forceinline__ void split8(const float (&v)[QW], float scale, uint32_t (&hi)[4], uint32_t (&lo)[4]) {
  const float2 sc2 = make_float2(scale, scale);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 s = __fmul2_rn(make_float2(v[2 * i], v[2 * i + 1]), sc2);
    const __half2 h = __float22half2_rn(s);
    const float2 hf = __half22float2(h);
    const __half2 l = __float22half2_rn(__fadd2_rn(s, make_float2(-hf.x, -hf.y)));
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
  uint64_t* bar_a;       // [2]  A chunks 0..2 / 3..4 written by a layer epilogue
  uint64_t* bar_in;      // [2]  A chunks 0..2 / 3..4 written by the tile's input producer
  uint64_t* bar_accfull; // [2]       accumulator part A / B complete
  uint64_t* bar_accempty;// [2]       accumulator part A / B read by all epilogue warps
  uint32_t* tmem_ptr;
};
constexpr int MAX_BIAS = 4;
constexpr size_t SMEM_BYTES = 2 * (2 * (size_t)IMG_BIG) + 2 * (size_t)FP * 32 * 2 + MAX_BIAS * FP * 4 + NQ * TILE * 4 + (3 * FP + 4) * 4 +
                              24 * 8 + 16 + 128 /* alignment slack */;
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
  s.bar_wsmall = b; s.bar_wfull = b + 1; s.bar_wempty = b + 3; s.bar_a = b + 5; s.bar_in = b + 10; s.bar_accfull = b + 15;
  s.bar_accempty = b + 17;
  s.tmem_ptr = reinterpret_cast<uint32_t*>(b + 20);
  return s;
}

// Program description (compile time).  in_src: where the layer's A comes from —
//   IN_PRODUCER  written by the tile's input producer      (wait bar_in[c])
//   IN_EPILOGUE  written by the previous layer's epilogue  (wait bar_a[c], A buffer toggles)
//   IN_SAME      the previous layer's A is reused          (no wait)
enum { IN_PRODUCER = 0, IN_EPILOGUE = 1, IN_SAME = 2 };
struct LayerStep { int layer; int ksteps; int in_src; };

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

// MMA issuer: one thread; per layer two column parts, each 3 MMAs per K step.
// Optional timeline capture (debug builds of the bench tool only): CTA 0's MMA thread records clock64 stamps.
#ifdef AGX_TC_TIMELINE
__device__ long long* g_tc_timeline = nullptr;
#define AGX_STAMP(slot) do { if (g_tc_timeline && blockIdx.x == 0 && stamp_i < 1024) g_tc_timeline[(NL == 4 ? 0 : NL == 6 ? 1024 : 2048) + stamp_i++] = ((long long)(slot) << 48) | (clock64() & 0xffffffffffffll); } while (0)
#else
#define AGX_STAMP(slot) do { } while (0)
#endif

template <int NL>
__device__ __forceinline__ void mma_role(const Shared& sh, const LayerStep (&prog)[NL], uint32_t tmem_base, int n_tiles) {
  if (!elect_one()) return;
  int stamp_i = 0; (void)stamp_i;
  uint32_t buf = 0, full_parity = 0, accempty_parity = 0x3, a_parity = 0, in_parity = 0;
  int a_cur = 0;   // A buffer the current layer reads
  bool small_ready = false;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      const int ksteps = prog[l].ksteps;
      const bool big = ksteps == 10;
      const int kpad = ksteps * 16;
      if (prog[l].in_src == IN_EPILOGUE) a_cur ^= 1;
      uint32_t w_addr;
      AGX_STAMP(1);
      if (big) {
        mbar_wait(&sh.bar_wfull[buf], (full_parity >> buf) & 1);
        full_parity ^= 1u << buf;
        w_addr = smem_u32(sh.wbig[buf]);
      } else {
        if (!small_ready) { mbar_wait(sh.bar_wsmall, 0); small_ready = true; }
        w_addr = smem_u32(sh.wsmall);
      }
      AGX_STAMP(2);
      const uint32_t img_bytes = (uint32_t)FP * kpad * 2;
      const uint32_t sbo = (uint32_t)(kpad >> 3) * 128;
      const uint32_t a_hi0 = tmem_base + col_ahi(a_cur), a_lo0 = tmem_base + col_alo(a_cur);
#pragma unroll
      for (int part = 0; part < 2; ++part) {
        mbar_wait(&sh.bar_accempty[part], (accempty_parity >> part) & 1);
        accempty_parity ^= 1u << part;
        AGX_STAMP(3 + part * 3);
        const uint32_t d_tmem = tmem_base + COL_ACC + (part ? N_PART_A : 0);
        const uint32_t row_off = part ? (uint32_t)(N_PART_A / 8) * sbo : 0u;
        const uint32_t idesc = part ? IDESC_B : IDESC_A;
        // descriptors advance by 256 B (= 16 in the encoded start address) per K step
        uint64_t b_hi = make_b_desc(w_addr + row_off, 128, sbo);
        uint64_t b_lo = make_b_desc(w_addr + img_bytes + row_off, 128, sbo);
        for (int ks = 0; ks < ksteps; ++ks) {
          if (part == 0 && (ks == 0 || ks == 2 * NCHUNK_A)) {
            const int h = ks ? 1 : 0;
            if (prog[l].in_src == IN_EPILOGUE) { mbar_wait(&sh.bar_a[h], (a_parity >> h) & 1); a_parity ^= 1u << h; }
            if (prog[l].in_src == IN_PRODUCER) { mbar_wait(&sh.bar_in[h], (in_parity >> h) & 1); in_parity ^= 1u << h; }
            tc_fence_after();
            AGX_STAMP(ks == 0 ? 4 : 5);
          }
          mma_f16_ts(d_tmem, a_lo0 + 8 * ks, b_hi, idesc, ks > 0);   // small terms first
          mma_f16_ts(d_tmem, a_hi0 + 8 * ks, b_lo, idesc, 1);
          mma_f16_ts(d_tmem, a_hi0 + 8 * ks, b_hi, idesc, 1);
          b_hi += 16;
          b_lo += 16;
        }
        mma_commit(&sh.bar_accfull[part]);
        AGX_STAMP(part ? 8 : 7);
      }
      if (big) { mma_commit(&sh.bar_wempty[buf]); buf ^= 1; }
    }
    a_cur ^= 1;   // the next tile's producer wrote the buffer the last layer was not reading
  }
}

// ------------------------------------------------------------------------------------------------ epilogue helpers
struct EpiCtx {
  int row;          // row of the tile this thread owns (= TMEM lane)
  int q;            // column quarter: this thread handles columns [32c + 8q, 32c + 8q + 8) of every chunk
  int lane, warp;
  uint32_t tmem_lane_base;   // tmem_base + (lane quarter << 16)
  uint32_t acc_parity;       // parity to wait for on bar_accfull[0/1] (both parts advance once per layer)
  int a_cur;        // A buffer the current layer reads (epilogues write the other one)
  int e_in;         // exponent of the scale applied to the current A
  float rowmax_in;  // max |a| of the current A row (unscaled)
};

// all of this warp's tensor-memory stores are complete and visible to the MMA warp -> one arrival per warp
__device__ __forceinline__ void epi_signal(const EpiCtx& cx, uint64_t* bar) {
  tmem_wait_st();
  tc_fence_before();
  __syncwarp();
  if (cx.lane == 0) mbar_arrive(bar);
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

// Writes this thread's 8 values of chunk c into A buffer `buf`.
__device__ __forceinline__ void epi_store_a(const EpiCtx& cx, int buf, int c, const float (&v)[QW], float scale) {
  uint32_t hi[4], lo[4];
  split8(v, scale, hi, lo);
  tmem_st4(cx.tmem_lane_base + col_ahi(buf) + 16 * c + 4 * cx.q, hi);
  tmem_st4(cx.tmem_lane_base + col_alo(buf) + 16 * c + 4 * cx.q, lo);
}

// One accumulator part (chunks CBEG..CEND-1) of a layer epilogue: wait for the part, read the thread's
// pieces, hand the part back to the MMA warp, then per chunk
//   v = acc * unscale (+ bias) ; extra(c, col0, v) ; [relu] ; consume(c, col0, v)
// and, if `signal` is set (the epilogue wrote the next layer's A), one mbarrier arrival per warp for the
// whole part after a single tcgen05.wait::st.
// UPFRONT: all pieces of the part are read first and the part is released before any arithmetic; otherwise
// (register-hungry epilogues) the pieces are read chunk by chunk and the part is released after the last read.
template <int PART, int CBEG, int CEND, bool RELU, bool UPFRONT, class Extra, class Consume>
__device__ __forceinline__ void epi_part(const Shared& sh, EpiCtx& cx, float unscale, const float* bias_s, uint64_t* signal, Extra& extra,
                                         Consume& consume) {
  mbar_wait(&sh.bar_accfull[PART], cx.acc_parity);
  tc_fence_after();
  const uint32_t acc = cx.tmem_lane_base + COL_ACC + QW * cx.q;
  const float2 us2 = make_float2(unscale, unscale);
  uint32_t r[UPFRONT ? CEND - CBEG : 1][QW];
  if (UPFRONT) {
#pragma unroll
    for (int c = CBEG; c < CEND; ++c) tmem_ld8(acc + 32 * c, r[c - CBEG]);
    tmem_wait_ld();
    tc_fence_before();
    __syncwarp();
    if (cx.lane == 0) mbar_arrive(&sh.bar_accempty[PART]);
  }
#pragma unroll
  for (int c = CBEG; c < CEND; ++c) {
    const int col0 = 32 * c + QW * cx.q;
    const int ri = UPFRONT ? c - CBEG : 0;
    if (!UPFRONT) {
      tmem_ld8(acc + 32 * c, r[0]);
      tmem_wait_ld();
      if (c == CEND - 1) {
        tc_fence_before();
        __syncwarp();
        if (cx.lane == 0) mbar_arrive(&sh.bar_accempty[PART]);
      }
    }
    float v[QW];
    float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
    if (bias_s) { b0 = lds128(bias_s + col0); b1 = lds128(bias_s + col0 + 4); }
    {
      const float2 p0 = __ffma2_rn(make_float2(__uint_as_float(r[ri][0]), __uint_as_float(r[ri][1])), us2, make_float2(b0.x, b0.y));
      const float2 p1 = __ffma2_rn(make_float2(__uint_as_float(r[ri][2]), __uint_as_float(r[ri][3])), us2, make_float2(b0.z, b0.w));
      const float2 p2 = __ffma2_rn(make_float2(__uint_as_float(r[ri][4]), __uint_as_float(r[ri][5])), us2, make_float2(b1.x, b1.y));
      const float2 p3 = __ffma2_rn(make_float2(__uint_as_float(r[ri][6]), __uint_as_float(r[ri][7])), us2, make_float2(b1.z, b1.w));
      v[0] = p0.x; v[1] = p0.y; v[2] = p1.x; v[3] = p1.y; v[4] = p2.x; v[5] = p2.y; v[6] = p3.x; v[7] = p3.y;
    }
    extra(c, col0, v);
    if (RELU) {
#pragma unroll
      for (int i = 0; i < QW; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    consume(c, col0, v);
  }
  if (signal) epi_signal(cx, signal);
}

// Layer epilogue over the two accumulator parts; a_out = true when consume() writes the next layer's A
// (arrivals on bar_a[0] after chunks 0..2 and on bar_a[1] after chunks 3..4).
template <bool RELU, bool UPFRONT = true, class Extra, class Consume>
__device__ __forceinline__ void epi_layer(const Shared& sh, EpiCtx& cx, float unscale, const float* bias_s, bool a_out, Extra extra,
                                          Consume consume) {
  epi_part<0, 0, NCHUNK_A, RELU, UPFRONT>(sh, cx, unscale, bias_s, a_out ? &sh.bar_a[0] : nullptr, extra, consume);
  epi_part<1, NCHUNK_A, NCHUNK, RELU, UPFRONT>(sh, cx, unscale, bias_s, a_out ? &sh.bar_a[1] : nullptr, extra, consume);
  cx.acc_parity ^= 1;
}

}  // namespace tc
}  // namespace agx
