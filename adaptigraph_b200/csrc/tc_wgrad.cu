// Weight gradients of the training path on the tensor cores (train.cu; reference: torch autograd of the nn.Linear layers in
// dynamics/gnn/model.py:4-60 under dynamics/train/train.py:90-112).
//
//   dW[n][j] += sum_m dY[m][n] * [mask[m][n] > 0] * X[m][j]          db[n] += sum_m dY[m][n] * [mask[m][n] > 0]
//
// is a GEMM whose contraction runs over the ROWS (relations or particles): D[160 x 160] = dY^T[160 x M] * X[M x 160].  Both operands
// are wanted "K-major" with K = the row index, i.e. transposed with respect to how the activations are stored, so the CTA
// transposes while it stages: the raw fp32 rows of a 32-row stage arrive by bulk copy (they are consecutive in memory; two stages
// in flight, no registers tied up), a thread then reads 8 consecutive rows x 4 columns and holds, per column, 8 consecutive K values
// of ONE operand row = exactly the 16 bytes of one core-matrix row of the no-swizzle K-major shared-memory layout (tc_ptx.cuh
// make_b_desc), written with a single st.shared.v4.  fp32 accuracy comes from the same split as
// the forward chains (x * 2^e = hi + lo in fp16, three products hi*hi + hi*lo + lo*hi, fp32 accumulation in tensor memory); because
// a sum over rows cannot be rescaled per row, one exact power-of-two scale per operand is shared by the rows a CTA accumulates: it
// is set from the first 64-row stage (maximum -> 2^12) and kept while later stages stay below 2^15; a stage that would not fit
// drains the accumulators into the partial product and restarts them under a new scale (rare).  An element keeps 22 significant
// bits down to 2^-14 of the scale's reference maximum and an absolute error of 2^-37 of it below, which is what a sum dominated
// by its large terms needs.  The constant 1 in the last padding column of X makes
// the bias gradient the last column of D.
//
// One launch serves a batch of independent jobs (every layer whose upstream gradient exists at that point of the backward), each
// split over `nctas` CTAs that leave fp32 partial products; wgrad_tc_reduce_kernel adds them in a fixed order (deterministic, no
// atomics) into the reference-layout gradient tensors.
#include "common.cuh"
#include "tc_chain.cuh"
#include "tc_wgrad.cuh"

namespace agx {
namespace tc {

constexpr int WG_ROWS = 32;                      // rows (= MMA K extent) per stage
constexpr int WG_THREADS = 512;
constexpr int WG_SBO = (WG_ROWS / 8) * 128 + 16; // bytes between 8-row groups of an operand image (+16: the transposing stores of
                                                 // lanes that differ in the 8-row group index fall into different banks)
constexpr int WG_ARR = (FP / 8) * WG_SBO;        // one operand image: 160 operand rows x 32 k, fp16
constexpr int WG_STAGE = 4 * WG_ARR;             // A_hi, A_lo, B_hi, B_lo
constexpr int WG_RAW_T = WG_ROWS * FP * 4;       // one raw fp32 tile: 32 rows x 160 columns
constexpr int WG_RAW = 3 * WG_RAW_T;             // dY, mask, X
constexpr int WG_OFF_RAW = 2 * WG_STAGE;
constexpr int WG_OFF_CTRL = WG_OFF_RAW + 2 * WG_RAW;
constexpr size_t WG_SMEM = (size_t)WG_OFF_CTRL + 512;
constexpr int WG_TMEM_COLS = 512;                // two accumulator blocks of 160 columns (block 1: operand rows 128..159)
constexpr int WG_SCALE_TARGET = 12;              // a fresh scale puts the stage maximum at <= 2^12 ...
constexpr float WG_SCALE_LIMIT = 32768.f;        // ... and is kept until a later stage would exceed 2^15 (fp16 overflows at 65504)
// The tensor core adds into its fp32 accumulator with truncation, a bias of about 2^-25 per MMA that grows linearly with the
// length of the chain (measured: weight gradients over 9632 rows off by 3e-5 of their maximum with one chain of 600 MMAs per
// CTA, 2e-6 with chains of 48).  The accumulators are therefore drained into the fp32 partial product (round-to-nearest adds)
// every WG_DRAIN_EVERY stages.
constexpr int WG_DRAIN_EVERY = 16;
static_assert(WG_SMEM <= 227 * 1024, "shared memory");

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// exponent e with bound * 2^e <= 2^WG_SCALE_TARGET
__device__ __forceinline__ int wg_scale_exp(float bound) { return scale_exp(bound) - (TARGET_EXP - WG_SCALE_TARGET); }

// The split of a batch over the grid is computed ON THE DEVICE, identically by every CTA and by the reduction: the number of
// relation rows that exist is only known there (AgxGraphIn.E_cap is a capacity, often twice the relations actually built), and a
// split by capacity would leave half the relation CTAs without work.  Every CTA gets the same number of stages, the smallest
// that fits the batch into `budget` CTAs; the assignment is a pure function of the job sizes (deterministic).
struct WgPlan {
  int spc;                          // stages per CTA
  int cta0[WG_MAX_JOBS + 1];        // first CTA of every job (cta0[njobs] = CTAs in use)
  int64_t M[WG_MAX_JOBS];           // rows that exist
};
__device__ __forceinline__ void wg_plan(const WgArgs& a, int budget, WgPlan& p) {
  int64_t total = 0;
  for (int i = 0; i < a.njobs; ++i) {
    const WgJob& j = a.job[i];
    p.M[i] = j.m_limit ? min(j.M, (int64_t)__ldg(j.m_limit)) : j.M;
    total += (p.M[i] + WG_ROWS - 1) / WG_ROWS;
  }
  int64_t spc = max((total + budget - 1) / budget, (int64_t)1);
  for (;; ++spc) {
    int64_t n = 0;
    for (int i = 0; i < a.njobs; ++i) n += ((p.M[i] + WG_ROWS - 1) / WG_ROWS + spc - 1) / spc;
    if (n <= budget) break;
  }
  p.spc = (int)spc;
  int c = 0;
  for (int i = 0; i < a.njobs; ++i) {
    p.cta0[i] = c;
    c += (int)(((p.M[i] + WG_ROWS - 1) / WG_ROWS + spc - 1) / spc);
  }
  p.cta0[a.njobs] = c;
}

__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tc_kernel(const WgArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar_done = reinterpret_cast<uint64_t*>(smem + WG_OFF_CTRL);         // [2] MMAs that read operand buffer b are complete
  uint64_t* bar_raw = bar_done + 2;                                             // [2] raw tile b has landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + WG_OFF_CTRL + 32);
  float* red = reinterpret_cast<float*>(smem + WG_OFF_CTRL + 64);               // [2 passes][2][16] per-warp maxima
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  int ji = 0;
  int64_t r0, r1;
  {
    WgPlan plan;
    wg_plan(a, (int)gridDim.x, plan);
    if ((int)blockIdx.x >= plan.cta0[a.njobs]) return;   // (uniform over the CTA) more CTAs than the rows that exist need
    while ((int)blockIdx.x >= plan.cta0[ji + 1]) ++ji;
    r0 = (int64_t)((int)blockIdx.x - plan.cta0[ji]) * plan.spc * WG_ROWS;
    r1 = min(plan.M[ji], r0 + (int64_t)plan.spc * WG_ROWS);
  }
  const WgJob& jb = a.job[ji];
  float* out = a.part + (size_t)blockIdx.x * (FP * FP);
  const int npad = jb.npad;
  const int nst = (int)((r1 - r0 + WG_ROWS - 1) / WG_ROWS);
  const bool has_mask = jb.mask != nullptr;
  const int ldx = jb.ldx;

  if (warp == 0) {
    tmem_alloc(tmem_slot, WG_TMEM_COLS);
    tc_fence_before();
  }
  if (tid == 0) {
    mbar_init(&bar_done[0], 1);
    mbar_init(&bar_done[1], 1);
    mbar_init(&bar_raw[0], 1);
    mbar_init(&bar_raw[1], 1);
    fence_mbar_init();
  }
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // raw fp32 rows of a stage are consecutive in memory: one bulk copy per tensor (TMA engine, no registers in flight)
  auto issue_raw = [&](int s) {   // issuer thread
    const int64_t m0 = r0 + (int64_t)s * WG_ROWS;
    const uint32_t rows = (uint32_t)min((int64_t)WG_ROWS, r1 - m0);
    uint8_t* raw = smem + WG_OFF_RAW + (s & 1) * WG_RAW;
    const uint32_t ba = rows * FP * 4, bx = rows * (uint32_t)ldx * 4;
    mbar_arrive_expect_tx(&bar_raw[s & 1], ba + (has_mask ? ba : 0u) + bx);
    bulk_g2s(raw, jb.dY + m0 * FP, ba, &bar_raw[s & 1]);
    if (has_mask) bulk_g2s(raw + WG_RAW_T, jb.mask + m0 * FP, ba, &bar_raw[s & 1]);
    bulk_g2s(raw + 2 * WG_RAW_T, jb.X + m0 * ldx, bx, &bar_raw[s & 1]);
  };
  if (tid == 15 * 32) {
    issue_raw(0);
    if (nst > 1) issue_raw(1);
  }

  const uint32_t idesc = make_idesc_f16(128, npad);
  int ncommit0 = 0, ncommit1 = 0, nwait0 = 0, nwait1 = 0;   // per operand buffer: MMA groups committed / observed complete
  auto wait_buf = [&](int buf) {
    if (buf) { while (nwait1 < ncommit1) { mbar_wait(&bar_done[1], nwait1 & 1); ++nwait1; } }
    else { while (nwait0 < ncommit0) { mbar_wait(&bar_done[0], nwait0 & 1); ++nwait0; } }
  };
  int ea = 0, eb = 0;
  float sa = 1.f, sb = 1.f;
  bool have_a = false, have_b = false, acc_valid = false, drained = false;

  // accumulators -> partial product (a store the first time, vector reductions afterwards).  warp w: lane quarter w % 4, 16-column chunks w / 4, + 4, ...
  auto drain = [&]() {
    wait_buf(0);
    wait_buf(1);
    tc_fence_after();
    const float unscale = exp2i(-ea - eb);
    const int q = warp & 3;
    for (int blk = 0; blk < 2; ++blk) {
      if (blk == 1 && q != 0) break;             // block 1: operand rows 128..159 = lanes 0..31
      const int n = blk * 128 + 32 * q + lane;
      for (int ch = warp >> 2; ch < npad / 16; ch += 4) {
        uint32_t r[16];
        tmem_ld16(tmem_base + ((uint32_t)(32 * q) << 16) + blk * FP + 16 * ch, r);
        tmem_wait_ld();
        float4* o = reinterpret_cast<float4*>(out + (size_t)n * FP + 16 * ch);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float4 v = make_float4(__uint_as_float(r[4 * i]) * unscale, __uint_as_float(r[4 * i + 1]) * unscale,
                                 __uint_as_float(r[4 * i + 2]) * unscale, __uint_as_float(r[4 * i + 3]) * unscale);
          if (drained)   // fire-and-forget vector add: no read round trip; one thread per address, in program order (deterministic)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + i), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
          else
            o[i] = v;
        }
      }
    }
    tc_fence_before();
    drained = true;
    acc_valid = false;
  };

  // conversion work, warps 0..14: an item is 8 rows x 2 columns.  Threads 0..319 convert one item of dY (with its mask), threads
  // 320..479 up to two items of X.  Lane 0 of warp 15 issues the bulk copies and the MMAs, so issuing overlaps the next conversion.
  const bool issuer = tid == 15 * 32;
  const int np2 = npad / 2;                      // column pairs of X that are staged (dY: FP / 2)
  const int n_x = (WG_ROWS / 8) * np2;
  // converts one item under `scale` into the operand image and returns its maximum
  auto convert_item = [&](bool is_a, int g, int pr, const uint8_t* raw, uint8_t* stage, int rows_here, float scale) -> float {
    const int c0 = 2 * pr;
    const bool colok = c0 < (is_a ? FP : jb.kx);
    const bool one = !is_a && jb.bias_col && c0 + 2 == npad;   // the pair whose last column carries the constant 1
    const int row_bytes = (is_a ? FP : ldx) * 4;
    const uint8_t* src = raw + (is_a ? 0 : 2 * WG_RAW_T) + (8 * g) * row_bytes + (colok ? 8 * pr : 0);
    float2 x[8];
    float mx = 0.f;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      x[m] = *reinterpret_cast<const float2*>(src + m * row_bytes);
      if (is_a && has_mask) {
        const float2 k = *reinterpret_cast<const float2*>(src + WG_RAW_T + m * row_bytes);
        x[m].x = k.x > 0.f ? x[m].x : 0.f;
        x[m].y = k.y > 0.f ? x[m].y : 0.f;
      }
      const bool in = 8 * g + m < rows_here;
      if (!(in && colok)) x[m] = make_float2(0.f, 0.f);
      if (one && in) x[m].y = 1.f;
      mx = fmaxf(mx, fmaxf(fabsf(x[m].x), fabsf(x[m].y)));
    }
    // scale, split, transpose: column c's 8 consecutive K values are one 16-byte core-matrix row
    uint8_t* img = stage + (is_a ? 0 : 2 * WG_ARR);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float e0 = j == 0 ? x[2 * k].x : x[2 * k].y, e1 = j == 0 ? x[2 * k + 1].x : x[2 * k + 1].y;
        const float2 sv = make_float2(e0 * scale, e1 * scale);
        const __half2 h = __float22half2_rn(sv);
        const float2 hf = __half22float2(h);
        const __half2 l = __float22half2_rn(make_float2(sv.x - hf.x, sv.y - hf.y));
        hi[k] = *reinterpret_cast<const uint32_t*>(&h);
        lo[k] = *reinterpret_cast<const uint32_t*>(&l);
      }
      const int col = c0 + j;
      uint8_t* dst = img + (col >> 3) * WG_SBO + g * 128 + (col & 7) * 16;
      *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(dst + WG_ARR) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
    return mx;
  };

  int pass = 0, since_drain = 0;
  uint32_t raw_par0 = 0, raw_par1 = 0;
  for (int s = 0; s < nst; ++s) {
    const int buf = s & 1;
    if (since_drain == WG_DRAIN_EVERY) { drain(); since_drain = 0; }
    ++since_drain;
    uint8_t* stage = smem + buf * WG_STAGE;
    const uint8_t* raw = smem + WG_OFF_RAW + buf * WG_RAW;
    const int64_t m0 = r0 + (int64_t)s * WG_ROWS;
    const int rows_here = (int)min((int64_t)WG_ROWS, r1 - m0);
#ifdef AGX_WG_TIMELINE
    const long long t0 = clock64();
#endif
    wait_buf(buf);   // the MMAs that last read this operand buffer are complete
#ifdef AGX_WG_TIMELINE
    const long long t1 = clock64();
#endif
    if (buf) { mbar_wait(&bar_raw[1], raw_par1); raw_par1 ^= 1; }
    else { mbar_wait(&bar_raw[0], raw_par0); raw_par0 ^= 1; }
#ifdef AGX_WG_TIMELINE
    const long long t2 = clock64();
#endif
    // A pass converts the stage under the current scales and finds its maxima on the way.  When a maximum does not fit (always in
    // the very first pass, where no scale exists yet) the accumulators are drained, the scale is renewed and the pass repeated.
    for (;;) {
      float mxa = 0.f, mxb = 0.f;
      if (tid < (WG_ROWS / 8) * (FP / 2)) {
        const int g = tid / (FP / 2);
        mxa = convert_item(true, g, tid - g * (FP / 2), raw, stage, rows_here, sa);
      } else if (tid < 480) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int xi = tid - 320 + 160 * r;
          if (xi < n_x) {
            const int g = xi / np2;
            mxb = fmaxf(mxb, convert_item(false, g, xi - g * np2, raw, stage, rows_here, sb));
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        mxa = fmaxf(mxa, __shfl_xor_sync(0xffffffffu, mxa, o));
        mxb = fmaxf(mxb, __shfl_xor_sync(0xffffffffu, mxb, o));
      }
      float* rd = red + 32 * (pass & 1);
      ++pass;
      if (lane == 0) { rd[warp] = mxa; rd[16 + warp] = mxb; }
      fence_proxy_async_smem();
      __syncthreads();
#pragma unroll
      for (int i = 0; i < WG_THREADS / 32; ++i) { mxa = fmaxf(mxa, rd[i]); mxb = fmaxf(mxb, rd[16 + i]); }
      // (an operand that has been all zero so far needs no scale: zeros are exact under any)
      const bool new_a = mxa > 0.f && (!have_a || mxa * sa > WG_SCALE_LIMIT), new_b = mxb > 0.f && (!have_b || mxb * sb > WG_SCALE_LIMIT);
      if (!(new_a || new_b)) break;   // uniform over the CTA
      if (acc_valid) drain();
      if (new_a) { ea = wg_scale_exp(mxa); sa = exp2i(ea); have_a = true; }
      if (new_b) { eb = wg_scale_exp(mxb); sb = exp2i(eb); have_b = true; }
    }
#ifdef AGX_WG_TIMELINE
    const long long t3 = clock64();
#endif
    if (issuer) {
      if (s + 2 < nst) issue_raw(s + 2);   // every thread has finished reading this raw buffer (barrier above)
      tc_fence_after();
      const uint32_t sa_hi = smem_u32(stage), sb_hi = sa_hi + 2 * WG_ARR;
#pragma unroll
      for (int blk = 0; blk < 2; ++blk) {
        const uint32_t d = tmem_base + blk * FP;
        const uint32_t a_off = blk * (128 / 8) * WG_SBO;
        uint64_t a_hi = make_b_desc(sa_hi + a_off, 128, WG_SBO), a_lo = make_b_desc(sa_hi + WG_ARR + a_off, 128, WG_SBO);
        uint64_t b_hi = make_b_desc(sb_hi, 128, WG_SBO), b_lo = make_b_desc(sb_hi + WG_ARR, 128, WG_SBO);
#pragma unroll
        for (int ks = 0; ks < WG_ROWS / 16; ++ks) {
          mma_f16_ss(d, a_lo, b_hi, idesc, (acc_valid || ks > 0) ? 1u : 0u);
          mma_f16_ss(d, a_hi, b_lo, idesc, 1u);
          mma_f16_ss(d, a_hi, b_hi, idesc, 1u);
          a_hi += 16; a_lo += 16; b_hi += 16; b_lo += 16;   // 256 bytes = two core matrices along K
        }
      }
      mma_commit(&bar_done[buf]);
#ifdef AGX_WG_TIMELINE
      if (blockIdx.x == 0 && s >= 4 && s < 24)
        printf("stage %d: wait_buf %lld wait_raw %lld convert %lld issue %lld passes %d\n", s, t1 - t0, t2 - t1, t3 - t2, clock64() - t3, pass);
#endif
    }
    if (buf) ++ncommit1; else ++ncommit0;
    acc_valid = true;
  }
  drain();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, WG_TMEM_COLS);
}

// dW[n * ld + col0 + j] += sum over the job's CTAs (in order) of part[cta][n][j]  (n < F, j < K);  db[n] += ... part[cta][n][npad-1].
// Jobs chained through `next` add into the same destination and are summed by the same thread, in chain order.
__global__ void wgrad_tc_reduce_kernel(const WgArgs a, const float* __restrict__ part, int budget) {
  const WgJob& head = a.job[blockIdx.y];
  if (!head.chain_head) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int F = head.F, K = head.K;
  if (i >= F * K + (head.db ? F : 0)) return;
  WgPlan plan;
  wg_plan(a, budget, plan);
  const bool is_b = i >= F * K;
  const int n = is_b ? i - F * K : i / K;
  const int j = is_b ? head.npad - 1 : i - n * K;
  float s = 0.f;
  for (int jj = blockIdx.y; jj >= 0; jj = a.job[jj].next) {
    const float* p = part + (size_t)plan.cta0[jj] * (FP * FP) + (size_t)n * FP + j;
    const int nctas = plan.cta0[jj + 1] - plan.cta0[jj];
    for (int c = 0; c < nctas; ++c) s += p[(size_t)c * (FP * FP)];
  }
  if (is_b) head.db[n] += s;
  else head.dW[(size_t)n * head.ld + head.col0 + j] += s;
}

}  // namespace tc

// ------------------------------------------------------------------------------------------------ host
size_t tc_wgrad_part_floats(int max_ctas) { return (size_t)max_ctas * FP * FP; }

int tc_wgrad_batch(cudaStream_t st, tc::WgArgs& a, float* part, int max_ctas) {
  using namespace tc;
  if (a.njobs <= 0) return AGX_OK;
  static thread_local DeviceOnce once;
  if (once.need()) AGX_CUDA_OK(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WG_SMEM));
  const int budget = max_ctas < num_sms() ? max_ctas : num_sms();   // one wave, one CTA per SM; the device plans the split
  a.part = part;
  { ProfScope ps(AGX_KIND_OTHER, st);
    wgrad_tc_kernel<<<budget, WG_THREADS, WG_SMEM, st>>>(a); }
  AGX_LAUNCH_CHECK();
  { ProfScope ps(AGX_KIND_OTHER, st);
    wgrad_tc_reduce_kernel<<<dim3((FP * FP + FP + 255) / 256, a.njobs), 256, 0, st>>>(a, part, budget); }
  AGX_LAUNCH_CHECK();
  return AGX_OK;
}

}  // namespace agx
