// Tensor-core MLP chains (AGX_PREC_TC_F16X3): tcgen05.mma kind::f16 with fp32-accurate split operands.
//
// One CTA keeps TWO tiles of 128 rows (relations or particles) in flight ("slots") and pushes each
// through a chain of dense layers without the activations ever leaving the SM.  The slots are
// independent pipelines that share the tensor pipe and the weights in shared memory, so one slot's
// epilogue (tensor-memory round trip, fp32 math, re-split) is hidden behind the other slot's MMAs.
//
//   * per slot, tensor memory holds the layer input A (160 K-elements as packed fp16: 80 hi + 80 lo
//     columns) and a 96-column fp32 accumulator: 256 columns per slot, 512 in total;
//   * every fp32 activation a is scaled by an exact per-row power of two and split a' = hi + lo
//     (22 significant bits); the weights are pre-packed (agx_pack_weights) as scaled fp16 hi/lo
//     images in the canonical K-major core-matrix layout and brought into shared memory by the TMA
//     engine as one bulk copy per layer, double buffered, shared by both slots;
//   * a layer is issued as two column parts (N = 96, then N = 64, same accumulator columns), each
//     3 MMAs per 16-wide K step:  D += Alo*Whi + Ahi*Wlo + Ahi*Whi  (dropped lo*lo term: 2^-22);
//   * biases ride in the MMA: the weight image carries bias[n] in the first padding column (k = 6 / 17 / 150) and every A row
//     carries a 1 there, so the epilogue of a plain layer is  s = acc * 2^(e_next - e_in - sw)  followed by the split, with
//     ReLU folded into the fp16 conversions (hi = cvt.rz.relu, lo = cvt.rn.relu of the residual);
//   * row scales come from propagated bounds (bound_out = bound_in * max_n(sum_k |W_nk| + |b_n|)), so a plain layer needs no
//     row maximum and no cross-warp exchange; actual row maxima are only taken where fp32 rows are stored anyway;
//   * 8 epilogue warps per slot (thread = row = tensor-memory lane; the two warps of a 32-lane
//     quarter split every 32-column chunk into 16-column halves) read an accumulator part with
//     tcgen05.ld, undo the power-of-two scales exactly, apply bias / residual / ReLU and either
//     re-split the result into the slot's next A (written once the layer's last MMA has completed:
//     part A's results wait in registers as packed fp16) or store fp32 rows to HBM (256-bit stores).
//
// Warp roles: warps 0-7 / 8-15 epilogue + input producers of slot 0 / 1, warps 16 / 17 MMA issuers of
// slot 0 / 1 (warp 16 also owns the TMEM allocation), warp 18 weight loader, warp 19 idle.
// All waits are bounded (tc_ptx.cuh: a protocol bug traps, it cannot hang).
#pragma once
#include "common.cuh"
#include "tc_ptx.cuh"

namespace agx {
namespace tc {

constexpr int TILE = 128;
constexpr int NSLOT = 2;
constexpr int HW = 16;                                  // columns per thread per 32-column chunk (two halves)
constexpr int SLOT_WARPS = 8;
constexpr int SLOT_THREADS = SLOT_WARPS * 32;
constexpr int EPI_WARPS = NSLOT * SLOT_WARPS;
constexpr int MMA_WARP0 = EPI_WARPS;                    // warps 16, 17
constexpr int LOAD_WARP = EPI_WARPS + 2;                // warp 18
constexpr int THREADS = (EPI_WARPS + 4) * 32;           // 20 warps
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t SLOT_COLS = 256;
constexpr uint32_t COL_ACC = 0, COL_AHI = 96, COL_ALO = 176;   // within a slot
constexpr int NCHUNK = 5;                               // 32-wide K chunks of a 160-wide layer
constexpr int NCHUNK_A = 3;                             // accumulator part A = chunks 0..2 (N = 96), part B = chunks 3..4 (N = 64)
constexpr int N_PART_A = 32 * NCHUNK_A, N_PART_B = FP - N_PART_A;
constexpr uint32_t IMG_BIG = FP * FP * 2;               // bytes of one fp16 image (hi or lo) of a 160x160 layer
constexpr uint32_t IDESC_A = make_idesc_f16(TILE, N_PART_A), IDESC_B = make_idesc_f16(TILE, N_PART_B);
constexpr int TARGET_EXP = 14;                          // scaled operands satisfy |x| <= 2^14 (fp16 max is 65504)

// tensor-core layer ids (order of the fp16 images in the packed blob)
enum TcLayer {
  T_PENC0 = 0, T_PENC2, T_PENC4, T_RENC0, T_RENC2, T_RENC4, T_RP_REL, T_RP_RECV, T_RP_SEND, T_PP_ENC, T_PP_AGG,
  T_PRED0, T_PRED1,
  // transposed images (element (j, n) = W[n][j], no bias column): the B operands of the backward's dX = dY * W products
  TT_PENC2, TT_PENC4, TT_RENC0, TT_RENC2, TT_RENC4, TT_RP_REL, TT_RP_RECV, TT_RP_SEND, TT_PP_ENC, TT_PP_AGG, TT_PRED0, TT_PRED1,
  T_NUM
};
__host__ __device__ constexpr int tc_kpad(int t) { return t == T_PENC0 ? 16 : (t == T_RENC0 ? 32 : FP); }

struct TcLayout {
  size_t meta;          // byte offset of float4 meta[T_NUM] = {2^-sw, max_n (sum_k |W_nk| + |b_n|), 0, 0}
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

// Blocked row storage of the tensor-core path.  Its intermediates ([rows][150] matrices: C, P, A_n, Qr, Qs and the split
// aggregate) live as [row / 128][piece][row % 128][width]: nine 16-column pieces and a tenth of 8 columns (144..151; the MMA's
// padding columns 152..159 are zero by construction and never stored — 152 instead of 160 columns is 5 % fewer bytes for every
// HBM-bound kernel).  An epilogue thread owns one row of a tile (= its tensor-memory lane) and 16 consecutive columns at a time,
// so the 32 lanes of a warp touch one contiguous 2 KB run instead of 32 rows 640 B apart (measured per SM: 30 instead of
// 20 B/clk for stores, 60 instead of 22 B/clk for loads).  Buffers are padded to whole tiles.
constexpr int BLK_W = 16;
constexpr int BLK_COLS = 152;                          // stored columns per row
constexpr int BLK_LAST = (BLK_COLS / BLK_W) * BLK_W;   // 144: first column of the narrow (8-column) piece
constexpr int BLK_TILE = TILE * BLK_COLS;              // floats per tile
__host__ __device__ constexpr int64_t blk_off(int64_t row, int col) {
  return (row >> 7) * (int64_t)BLK_TILE + (int64_t)(col >> 4) * (TILE * BLK_W) +
         (row & (TILE - 1)) * (col < BLK_LAST ? BLK_W : BLK_COLS - BLK_LAST) + (col & (BLK_W - 1));
}
__host__ __device__ constexpr int64_t blk_rows(int64_t rows) { return (rows + TILE - 1) / TILE * TILE; }

// exponent e such that bound * 2^e <= 2^TARGET_EXP (exact power-of-two scaling)
__device__ __forceinline__ int scale_exp(float bound) {
  if (!(bound > 0.f)) return 0;
  const int x = (int)((__float_as_uint(bound) >> 23) & 0xff) - 126;   // bound <= 2^x
  return max(-100, min(100, TARGET_EXP - x));
}
__device__ __forceinline__ float exp2i(int e) { return __uint_as_float((uint32_t)(e + 127) << 23); }

// packed conversions with ReLU folded in: {lo16: cvt(a), hi16: cvt(b)}
__device__ __forceinline__ uint32_t cvt_rz_relu_f16x2(float a, float b) {
  uint32_t r;
  asm("cvt.rz.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ uint32_t cvt_rn_relu_f16x2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
// relu(s) split into packed fp16 hi / lo columns.  hi truncates towards zero, so the residual of a positive value is
// non-negative and the second relu-conversion only clips what belongs to negative inputs (hi = 0, residual = s < 0).
__device__ __forceinline__ void split16_relu(const float (&s)[HW], uint32_t (&hi)[8], uint32_t (&lo)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t h = cvt_rz_relu_f16x2(s[2 * i], s[2 * i + 1]);
    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h));
    const float2 l = __fadd2_rn(make_float2(s[2 * i], s[2 * i + 1]), make_float2(-hf.x, -hf.y));
    hi[i] = h;
    lo[i] = cvt_rn_relu_f16x2(l.x, l.y);
  }
}
// hi-only variants (the consuming layer issues 2 MMAs per K step): round to nearest
__device__ __forceinline__ void round16_relu(const float (&s)[HW], uint32_t (&hi)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) hi[i] = cvt_rn_relu_f16x2(s[2 * i], s[2 * i + 1]);
}
__device__ __forceinline__ void round16(const float (&v)[HW], float scale, uint32_t (&hi)[8]) {
  const float2 sc2 = make_float2(scale, scale);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const __half2 h = __float22half2_rn(__fmul2_rn(make_float2(v[2 * i], v[2 * i + 1]), sc2));
    hi[i] = *reinterpret_cast<const uint32_t*>(&h);
  }
}
// signed variant for layer inputs that may be negative (hi = round-to-nearest, lo = residual)
__device__ __forceinline__ void split16(const float (&v)[HW], float scale, uint32_t (&hi)[8], uint32_t (&lo)[8]) {
  const float2 sc2 = make_float2(scale, scale);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float2 s = __fmul2_rn(make_float2(v[2 * i], v[2 * i + 1]), sc2);
    const __half2 h = __float22half2_rn(s);
    const float2 hf = __half22float2(h);
    const __half2 l = __float22half2_rn(__fadd2_rn(s, make_float2(-hf.x, -hf.y)));
    hi[i] = *reinterpret_cast<const uint32_t*>(&h);
    lo[i] = *reinterpret_cast<const uint32_t*>(&l);
  }
}
constexpr int ONE_COL = 150;   // A column that carries the constant 1 feeding the bias column of a 160-wide layer

// ------------------------------------------------------------------------------------------------
struct Shared {
  uint8_t* wbig[2];      // two 2*IMG_BIG weight buffers (hi image then lo image), shared by both slots
  uint8_t* wsmall;       // first-layer weights (K = 16 or 32)
  float* xchg;           // [NSLOT][2][TILE] row exchange between the two column halves of a slot
  float* head_w;         // [3][FP] + [4] (head program only)
  uint64_t* bar_wsmall;
  uint64_t* bar_wfull;   // [2]
  uint64_t* bar_wempty;  // [2]      count NSLOT: one arrival per slot per use
  uint64_t* bar_in;      // [NSLOT]  slot's A written by the tile's input producer
  uint64_t* bar_a;       // [NSLOT]  slot's A written by a layer epilogue
  uint64_t* bar_accfull; // [NSLOT]  accumulator part complete (parts alternate on the same barrier)
  uint64_t* bar_accempty;// [NSLOT]  accumulator part read by all epilogue warps of the slot
  uint32_t* tmem_ptr;
};
constexpr size_t SMEM_BYTES = 2 * (2 * (size_t)IMG_BIG) + 2 * (size_t)FP * 32 * 2 + NSLOT * 2 * TILE * 4 +
                              (3 * FP + 4) * 4 + 16 * 8 + 16 + 128 /* alignment slack */;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB shared-memory limit of a CTA");

__device__ __forceinline__ Shared carve_shared(uint8_t* raw) {
  uint8_t* p = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 127) & ~(uintptr_t)127);
  Shared s;
  s.wbig[0] = p; p += 2 * IMG_BIG;
  s.wbig[1] = p; p += 2 * IMG_BIG;
  s.wsmall = p; p += 2 * FP * 32 * 2;
  s.xchg = reinterpret_cast<float*>(p); p += NSLOT * 2 * TILE * 4;
  s.head_w = reinterpret_cast<float*>(p); p += (3 * FP + 4) * 4;
  uint64_t* b = reinterpret_cast<uint64_t*>(p);
  s.bar_wsmall = b; s.bar_wfull = b + 1; s.bar_wempty = b + 3; s.bar_in = b + 5; s.bar_a = b + 7; s.bar_accfull = b + 9;
  s.bar_accempty = b + 11;
  s.tmem_ptr = reinterpret_cast<uint32_t*>(b + 13);
  return s;
}

// Program description (compile time).  in_src: where the layer's A comes from —
//   IN_PRODUCER  written by the tile's input producer      (wait bar_in)
//   IN_EPILOGUE  written by the previous layer's epilogue  (wait bar_a)
//   IN_SAME      the previous layer's A is reused          (no wait)
enum { IN_PRODUCER = 0, IN_EPILOGUE = 1, IN_SAME = 2 };
// nmma: partial products issued per K step (the precision budget of the layer, tests/bench/precision_study.py):
//   3  Alo*Whi + Ahi*Wlo + Ahi*Whi   A and W at 22 significant bits (fp32-accurate)
//   2  Ahi*Wlo + Ahi*Whi             A rounded to fp16 (11 bits, a random per-element error), W at 22 bits; the layer's A carries
//                                    no lo columns: the producing epilogue neither computes nor stores them
struct LayerStep { int layer; int ksteps; int in_src; int nmma; };
// does the A operand read by layer l need its lo columns?  (layers marked IN_SAME share the A of the layer before them)
template <int NL>
__host__ __device__ constexpr bool a_needs_lo(const LayerStep (&prog)[NL], int l) {
  bool need = prog[l].nmma == 3;
  for (int i = l + 1; i < NL && prog[i].in_src == IN_SAME; ++i) need = need || prog[i].nmma == 3;
  return need;
}

// tiles of this CTA: t(k) = blockIdx.x + k * gridDim.x; slot s owns k = s, s + 2, ...; a "round" is one tile per slot
__device__ __forceinline__ int cta_tile_count(int n_tiles) {
  return (int)blockIdx.x < n_tiles ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
}

// Optional timeline capture (tools/tc_timeline.py builds with -DAGX_TC_TIMELINE=<kernel id>): in that kernel CTA 0
// records clock64 stamps of its two MMA threads (regions 0, 1), the weight loader (2) and lane 0 of epilogue warps 0 / 4 (3, 4:
// the two column halves of slot 0).  Region r holds 800 stamps at [r * 800, ...): (id << 48) | clock.
#ifdef AGX_TC_TIMELINE
__device__ long long* g_tc_timeline = nullptr;
__device__ __forceinline__ void tl_stamp(int region, int& i, int id) {
  if (g_tc_timeline && blockIdx.x == 0 && i < 800) g_tc_timeline[region * 800 + i++] = ((long long)id << 48) | (clock64() & 0xffffffffffffll);
}
#define AGX_STAMP(id) do { if (kid == AGX_TC_TIMELINE) tl_stamp(slot, stamp_i, id); } while (0)
#define AGX_STAMP_LOADER(id) do { if (kid == AGX_TC_TIMELINE) tl_stamp(2, stamp_i, id); } while (0)
#define AGX_STAMP_EPI(cx, id) do { if ((cx).tl_region >= 0) tl_stamp((cx).tl_region, (cx).tl_i, id); } while (0)
#else
#define AGX_STAMP(id) do { } while (0)
#define AGX_STAMP_LOADER(id) do { } while (0)
#define AGX_STAMP_EPI(cx, id) do { } while (0)
#endif

// ------------------------------------------------------------------------------------------------ roles
// Weight loader: one thread streams the big layers of every round through the two-buffer ring.
template <int NL>
__device__ __forceinline__ void loader_role(const Shared& sh, const LayerStep (&prog)[NL], const uint8_t* blob, const TcLayout& L,
                                            int n_tiles, int kid, bool keep_weights = false) {
  (void)kid;
  if (!elect_one()) return;
  const int my_tiles = cta_tile_count(n_tiles);
  if (my_tiles == 0) return;
  // keep_weights: mark the weight images evict_last in L2.  Pays where the kernel streams about a gigabyte through L2 per launch
  // and re-reads 300 KB of weights for every pair of tiles (node update: 0.229 -> 0.216 ms, its DRAM reads were 0.12 GB above the
  // algorithmic bytes); costs ~1 % in the encoder chains, which stream little and never lose their weights.
  const uint64_t keep = keep_weights ? l2_policy_evict_last() : 0;
  auto copy_w = [&](void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    if (keep_weights) bulk_g2s_hint(dst, src, bytes, bar, keep);
    else bulk_g2s(dst, src, bytes, bar);
  };
  {
    const int t = prog[0].layer;
    if (prog[0].ksteps < 10) {
      const uint32_t bytes = 2u * FP * tc_kpad(t) * 2u;
      mbar_arrive_expect_tx(sh.bar_wsmall, bytes);
      copy_w(sh.wsmall, blob + L.img[t], bytes, sh.bar_wsmall);
    }
  }
  uint32_t buf = 0, empty_parity = 0x3;   // bit b = parity to wait for on bar_wempty[b] (starts at 1: free)
  int stamp_i = 0; (void)stamp_i;
  const int rounds = (my_tiles + NSLOT - 1) / NSLOT;
  for (int round = 0; round < rounds; ++round) {
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      if (prog[l].ksteps < 10) continue;
      AGX_STAMP_LOADER(10);
      mbar_wait(&sh.bar_wempty[buf], (empty_parity >> buf) & 1);
      AGX_STAMP_LOADER(11);
      empty_parity ^= 1u << buf;
      mbar_arrive_expect_tx(&sh.bar_wfull[buf], 2 * IMG_BIG);
      copy_w(sh.wbig[buf], blob + L.img[prog[l].layer], 2 * IMG_BIG, &sh.bar_wfull[buf]);
      buf ^= 1;
    }
  }
}

// MMA issuer of one slot: one thread; per layer two column parts, each 3 MMAs per K step.
template <int NL>
__device__ __forceinline__ void mma_role(const Shared& sh, const LayerStep (&prog)[NL], int slot, uint32_t tmem_base, int n_tiles,
                                         int kid) {
  (void)kid;
  if (!elect_one()) return;
  int stamp_i = 0; (void)stamp_i;
  const int my_tiles = cta_tile_count(n_tiles);
  const int rounds = (my_tiles + NSLOT - 1) / NSLOT;
  const uint32_t tslot = tmem_base + slot * SLOT_COLS;
  const uint32_t d_tmem = tslot + COL_ACC, a_hi0 = tslot + COL_AHI, a_lo0 = tslot + COL_ALO;
  uint32_t buf = 0, full_parity = 0, accempty_parity = 1, a_parity = 0, in_parity = 0;
  bool small_ready = false;
  for (int round = 0; round < rounds; ++round) {
    const bool active = round * NSLOT + slot < my_tiles;
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      const int ksteps = prog[l].ksteps;
      const bool big = ksteps == 10;
      const int kpad = ksteps * 16;
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
      if (active) {
        const uint32_t img_bytes = (uint32_t)FP * kpad * 2;
        const uint32_t sbo = (uint32_t)(kpad >> 3) * 128;
#pragma unroll
        for (int part = 0; part < 2; ++part) {
          mbar_wait(&sh.bar_accempty[slot], accempty_parity);
          accempty_parity ^= 1;
          AGX_STAMP(3 + part * 3);
          if (part == 0) {
            if (prog[l].in_src == IN_EPILOGUE) { mbar_wait(&sh.bar_a[slot], a_parity); a_parity ^= 1; }
            if (prog[l].in_src == IN_PRODUCER) { mbar_wait(&sh.bar_in[slot], in_parity); in_parity ^= 1; }
            AGX_STAMP(4);
          }
          tc_fence_after();
          const uint32_t row_off = part ? (uint32_t)(N_PART_A / 8) * sbo : 0u;
          const uint32_t idesc = part ? IDESC_B : IDESC_A;
          // descriptors advance by 256 B (= 16 in the encoded start address) per K step
          uint64_t b_hi = make_b_desc(w_addr + row_off, 128, sbo);
          uint64_t b_lo = make_b_desc(w_addr + img_bytes + row_off, 128, sbo);
          const bool three = prog[l].nmma == 3;
          for (int ks = 0; ks < ksteps; ++ks) {
            if (three) mma_f16_ts(d_tmem, a_lo0 + 8 * ks, b_hi, idesc, ks > 0);   // small terms first
            mma_f16_ts(d_tmem, a_hi0 + 8 * ks, b_lo, idesc, three || ks > 0);
            mma_f16_ts(d_tmem, a_hi0 + 8 * ks, b_hi, idesc, 1);
            b_hi += 16;
            b_lo += 16;
          }
          mma_commit(&sh.bar_accfull[slot]);
          AGX_STAMP(part ? 8 : 7);
        }
        if (big) mma_commit(&sh.bar_wempty[buf]);
      } else if (big) {
        mbar_arrive(&sh.bar_wempty[buf]);   // keep the weight ring in lock step when this slot has no tile in the round
      }
      if (big) buf ^= 1;
    }
  }
}

// ------------------------------------------------------------------------------------------------ epilogue helpers
struct EpiCtx {
  int slot;         // which in-flight tile this warp serves
  int row;          // row of the tile this thread owns (= TMEM lane)
  int half;         // this thread handles columns [32c + 16*half, +16) of every chunk
  int lane, warp;
  uint32_t tslot;   // tmem_base + slot*SLOT_COLS + (lane quarter << 16)
  uint32_t acc_parity;   // parity to wait for on bar_accfull[slot] (flips after every part)
  int e_in;         // exponent of the scale applied to the current A
  float bound_in;   // upper bound on max |a| of the current A row (unscaled; >= 1 when the row carries the bias 1)
#ifdef AGX_TC_TIMELINE
  int tl_region, tl_i;
#endif
};

// all of this warp's tensor-memory stores are complete and visible to the MMA warp -> one arrival per warp
__device__ __forceinline__ void epi_signal(const EpiCtx& cx, uint64_t* bar) {
  tmem_wait_st();
  tc_fence_before();
  __syncwarp();
  if (cx.lane == 0) mbar_arrive(bar);
}

// combine a per-thread partial row value across the two column halves of the slot (max or sum)
template <bool IS_MAX>
__device__ __forceinline__ float epi_exchange(const Shared& sh, const EpiCtx& cx, float v) {
  float* x = sh.xchg + cx.slot * 2 * TILE;
  sts32(&x[cx.half * TILE + cx.row], v);
  named_bar_sync(1 + cx.slot, SLOT_THREADS);
  const float a = lds32(&x[cx.row]), b = lds32(&x[TILE + cx.row]);   // same order in both threads: identical result
  named_bar_sync(1 + cx.slot, SLOT_THREADS);
  return IS_MAX ? fmaxf(a, b) : a + b;
}

// Writes this thread's 16 values of chunk c into the slot's A.
__device__ __forceinline__ void epi_store_packed(const EpiCtx& cx, int c, const uint32_t (&hi)[8], const uint32_t (&lo)[8]) {
  tmem_st8(cx.tslot + COL_AHI + 16 * c + 8 * cx.half, hi);
  tmem_st8(cx.tslot + COL_ALO + 16 * c + 8 * cx.half, lo);
}
__device__ __forceinline__ void epi_store_hi(const EpiCtx& cx, int c, const uint32_t (&hi)[8]) {
  tmem_st8(cx.tslot + COL_AHI + 16 * c + 8 * cx.half, hi);
}
template <bool LO = true>
__device__ __forceinline__ void epi_store_a(const EpiCtx& cx, int c, const float (&v)[HW], float scale) {
  uint32_t hi[8], lo[8];
  if (LO) {
    split16(v, scale, hi, lo);
    epi_store_packed(cx, c, hi, lo);
  } else {
    round16(v, scale, hi);
    epi_store_hi(cx, c, hi);
  }
}

// wait for the next accumulator part of the slot
__device__ __forceinline__ void epi_wait_part(const Shared& sh, EpiCtx& cx) {
  mbar_wait(&sh.bar_accfull[cx.slot], cx.acc_parity);
  cx.acc_parity ^= 1;
  tc_fence_after();
}
// every thread of the warp has its pieces of the part in registers -> hand the accumulator back
__device__ __forceinline__ void epi_release_part(const Shared& sh, const EpiCtx& cx) {
  tc_fence_before();
  __syncwarp();
  if (cx.lane == 0) mbar_arrive(&sh.bar_accempty[cx.slot]);
}

// v = acc * f for one 16-column piece
__device__ __forceinline__ void epi_scale(const uint32_t (&r)[HW], float f, float (&v)[HW]) {
  const float2 f2 = make_float2(f, f);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float2 p = __fmul2_rn(make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), f2);
    v[2 * i] = p.x; v[2 * i + 1] = p.y;
  }
}
// the thread that owns A column ONE_COL (chunk 4, upper half, element 6)
__device__ __forceinline__ bool owns_one(const EpiCtx& cx, int c) { return c == ONE_COL / 32 && cx.half == (ONE_COL % 32) / HW; }

// Plain hidden layer: next A = relu(acc * 2^-(e_in + sw))  (the bias is already inside acc), rescaled by `scale` and split.
// Part A's packed results wait in registers until the layer's last MMA (part B) has completed — only then may A be overwritten.
template <bool LO = true>
__device__ __forceinline__ void epi_layer_plain(const Shared& sh, EpiCtx& cx, float unscale, float scale) {
  const float f = unscale * scale;
  uint32_t hiA[NCHUNK_A][8], loA[LO ? NCHUNK_A : 1][8];
  AGX_STAMP_EPI(cx, 20);
  epi_wait_part(sh, cx);
  AGX_STAMP_EPI(cx, 21);
  {
    uint32_t r[NCHUNK_A][HW];
#pragma unroll
    for (int c = 0; c < NCHUNK_A; ++c) tmem_ld16(cx.tslot + COL_ACC + 32 * c + HW * cx.half, r[c]);
    tmem_wait_ld();
    epi_release_part(sh, cx);
    AGX_STAMP_EPI(cx, 22);
#pragma unroll
    for (int c = 0; c < NCHUNK_A; ++c) {
      float v[HW];
      epi_scale(r[c], f, v);
      if (LO) split16_relu(v, hiA[c], loA[LO ? c : 0]);
      else round16_relu(v, hiA[c]);
    }
  }
  AGX_STAMP_EPI(cx, 23);
  epi_wait_part(sh, cx);   // part B complete => every MMA of the layer has read A
  AGX_STAMP_EPI(cx, 24);
#pragma unroll
  for (int c = 0; c < NCHUNK_A; ++c) {
    if (LO) epi_store_packed(cx, c, hiA[c], loA[LO ? c : 0]);
    else epi_store_hi(cx, c, hiA[c]);
  }
#pragma unroll
  for (int c = NCHUNK_A; c < NCHUNK; ++c) {
    uint32_t r[HW];
    tmem_ld16(cx.tslot + COL_ACC + 32 * (c - NCHUNK_A) + HW * cx.half, r);
    tmem_wait_ld();
    if (c == NCHUNK - 1) epi_release_part(sh, cx);
    float v[HW];
    epi_scale(r, f, v);
    if (owns_one(cx, c)) v[ONE_COL % HW] = scale;   // the 1 that multiplies the next layer's bias column
    uint32_t hi[8], lo[8];
    if (LO) {
      split16_relu(v, hi, lo);
      epi_store_packed(cx, c, hi, lo);
    } else {
      round16_relu(v, hi);
      epi_store_hi(cx, c, hi);
    }
  }
  AGX_STAMP_EPI(cx, 25);
  epi_signal(cx, &sh.bar_a[cx.slot]);
  AGX_STAMP_EPI(cx, 26);
}

// General layer whose result becomes the slot's next A.  per chunk: v = acc * unscale ; extra(c, col0, v) ; relu (RELU) ;
// side(c, col0, v) [e.g. store the fp32 row piece; it may also replace v: what it leaves is what gets split] ; split with `scale`.
// Returns the thread's partial row maximum of |v| (taken before side).
// (RELU = false is the backward use: no constant-1 column either -- the transposed images have no bias column, and a 1 times the
// large scale of a small gradient row would overflow fp16.)
template <bool LO = true, bool RELU = true, class Extra, class Side>
__device__ __forceinline__ float epi_layer_to_a(const Shared& sh, EpiCtx& cx, float unscale, float scale, Extra extra, Side side) {
  float mx = 0.f;
  uint32_t hiA[NCHUNK_A][8], loA[LO ? NCHUNK_A : 1][8];
  AGX_STAMP_EPI(cx, 40);
  epi_wait_part(sh, cx);
  AGX_STAMP_EPI(cx, 41);
  {
    uint32_t r[NCHUNK_A][HW];
#pragma unroll
    for (int c = 0; c < NCHUNK_A; ++c) tmem_ld16(cx.tslot + COL_ACC + 32 * c + HW * cx.half, r[c]);
    tmem_wait_ld();
    epi_release_part(sh, cx);
#pragma unroll
    for (int c = 0; c < NCHUNK_A; ++c) {
      const int col0 = 32 * c + HW * cx.half;
      float v[HW];
      epi_scale(r[c], unscale, v);
      extra(c, col0, v);
#pragma unroll
      for (int i = 0; i < HW; ++i) { if (RELU) v[i] = fmaxf(v[i], 0.f); mx = fmaxf(mx, fabsf(v[i])); }
      side(c, col0, v);
      if (LO) split16(v, scale, hiA[c], loA[LO ? c : 0]);
      else round16(v, scale, hiA[c]);
    }
  }
  AGX_STAMP_EPI(cx, 43);
  epi_wait_part(sh, cx);   // part B complete => every MMA of the layer has read A
  AGX_STAMP_EPI(cx, 44);
#pragma unroll
  for (int c = 0; c < NCHUNK_A; ++c) {
    if (LO) epi_store_packed(cx, c, hiA[c], loA[LO ? c : 0]);
    else epi_store_hi(cx, c, hiA[c]);
  }
#pragma unroll
  for (int c = NCHUNK_A; c < NCHUNK; ++c) {
    const int col0 = 32 * c + HW * cx.half;
    uint32_t r[HW];
    tmem_ld16(cx.tslot + COL_ACC + 32 * (c - NCHUNK_A) + HW * cx.half, r);
    tmem_wait_ld();
    if (c == NCHUNK - 1) epi_release_part(sh, cx);
    float v[HW];
    epi_scale(r, unscale, v);
    extra(c, col0, v);
#pragma unroll
    for (int i = 0; i < HW; ++i) { if (RELU) v[i] = fmaxf(v[i], 0.f); mx = fmaxf(mx, fabsf(v[i])); }
    side(c, col0, v);
    if (RELU && owns_one(cx, c)) v[ONE_COL % HW] = 1.f;   // after the fp32 side store: only the tensor-memory copy carries the 1
    epi_store_a<LO>(cx, c, v, scale);
  }
  AGX_STAMP_EPI(cx, 45);
  epi_signal(cx, &sh.bar_a[cx.slot]);
  return mx;
}

// Layer epilogue that only consumes the result (fp32 rows to HBM, running dot products ...): per chunk
// v = acc * unscale ; [relu] ; consume(c, col0, v).  A is left untouched by the layer itself; all_read() runs as soon as the
// layer's last MMA has completed, i.e. when the slot's A may be overwritten (the next tile's input is committed there).
// Each part is pulled into registers and handed back to the MMA warp before its (slow) consumers run.
struct NoHook { __device__ void operator()() const {} };
template <bool RELU, class Consume, class AllRead = NoHook>
__device__ __forceinline__ void epi_layer_out(const Shared& sh, EpiCtx& cx, float unscale, Consume consume, AllRead all_read = AllRead{}) {
  auto finish = [&](int c, const uint32_t (&r)[HW]) {
    float v[HW];
    epi_scale(r, unscale, v);
    if (RELU) {
#pragma unroll
      for (int i = 0; i < HW; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    consume(c, 32 * c + HW * cx.half, v);
  };
  AGX_STAMP_EPI(cx, 32);
  epi_wait_part(sh, cx);
  AGX_STAMP_EPI(cx, 34);
  {
    uint32_t r[NCHUNK_A][HW];
#pragma unroll
    for (int c = 0; c < NCHUNK_A; ++c) tmem_ld16(cx.tslot + COL_ACC + 32 * c + HW * cx.half, r[c]);
    tmem_wait_ld();
    epi_release_part(sh, cx);
    AGX_STAMP_EPI(cx, 35);
#pragma unroll
    for (int c = 0; c < NCHUNK_A; ++c) finish(c, r[c]);
  }
  AGX_STAMP_EPI(cx, 36);
  epi_wait_part(sh, cx);
  AGX_STAMP_EPI(cx, 37);
  {
    uint32_t r[NCHUNK - NCHUNK_A][HW];
#pragma unroll
    for (int c = NCHUNK_A; c < NCHUNK; ++c) tmem_ld16(cx.tslot + COL_ACC + 32 * (c - NCHUNK_A) + HW * cx.half, r[c - NCHUNK_A]);
    tmem_wait_ld();
    epi_release_part(sh, cx);
    all_read();
    AGX_STAMP_EPI(cx, 38);
#pragma unroll
    for (int c = NCHUNK_A; c < NCHUNK; ++c) finish(c, r[c - NCHUNK_A]);
  }
  AGX_STAMP_EPI(cx, 33);
}

}  // namespace tc
}  // namespace agx
