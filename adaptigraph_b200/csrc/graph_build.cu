// Radius AND top-k directed relation builder emitting CSR-by-receiver edge lists.
//
// Replaces construct_edges_from_states_batch (reference dynamics/dataset/graph.py:91-156)
// and construct_edges_from_states (:38-89).  The reference materialises B x N x N x 3
// difference tensors, a B x N x N distance matrix, a top-k index tensor and two dense
// B x n_rel x N one-hots; here nothing quadratic ever reaches HBM:
//
//   (G0 tool list    with connect_tools_all: ascending list of tool particles per graph, built by sort_cells' CTA)
//   G0b sort_cells   per graph: lay a uniform grid of cells of HALF the radius (GRID_SUBDIV = 2; wider once an axis would need
//                    more than 64 cells) over the two coordinate axes of largest extent, counting-sort the particles by
//                    cell id (shared-memory histogram + block scan), emit the permuted SoA copy of the graph and the
//                    first slot of every cell
//   G1 knn_rows      one thread per receiver: a sender in range lies within R cells of the receiver's cell along both
//                    grid axes (R = 2 normally).  The thread walks the square rings m = 0 .. R around its cell -- each a
//                    few contiguous slot runs of the sorted SoA copy (L1-resident: the threads of a CTA are a few
//                    neighbouring cells) -- keeping the k nearest in-radius senders in a per-thread sorted list in
//                    shared memory, and stops after ring m as soon as the list is full and its k-th distance is below
//                    m cell widths: every sender not yet seen is at least that far away.  With a small top-k inside a
//                    populous radius (cloth: 5 of ~28) that is a quarter of the candidates of the full window.  The
//                    list is finally re-sorted by sender id -> <= k candidates
//   G2 degrees_scan  per-row relation count after the tool rules (:134-144 / :77-80) + block scan; the last block to
//                    finish scans the block sums -> row offsets, total
//   G3 fill_rows     one thread per receiver merges candidates and tool senders in ascending sender order
//
// Arithmetic follows the reference exactly: dis = (dx*dx + dy*dy) + dz*dz in fp32 without FMA
// contraction (:109-110), radius test (dis - thr^2) < 0 (:125), pairs with an invalid endpoint or
// tool-tool pairs excluded (:111-118).  Distance ties are broken towards the lower sender id
// (torch.topk leaves the order of ties unspecified).
#include "common.cuh"

namespace agx {

constexpr int G1_THREADS = 128;
constexpr int G1_ROWS_PER_CTA = G1_THREADS;   // one thread per receiver
constexpr int SCAN_BLOCK = 1024;
constexpr unsigned FULL = 0xffffffffu;
constexpr int GRID_MAX_AXIS = 64;                                  // cells per grid axis (cells widen beyond the radius past that)
constexpr int GRID_MAX_CELLS = GRID_MAX_AXIS * GRID_MAX_AXIS;
constexpr int GRID_SUBDIV = 2;                                     // cells per radius (rings searched: up to GRID_SUBDIV)

struct GraphWs {
  int32_t* cand;       // [B*N][topk]
  int32_t* cnt;        // [B*N]   cnt | (non-tool cnt << 8)
  int32_t* deg;        // [B*N]
  int32_t* lpre;       // [B*N]   block-local exclusive prefix
  int32_t* blk;        // [nblk]  block sums -> exclusive block offsets
  int32_t* total;      // [1]
  int32_t* ticket;     // [1]     last-block-done counter of degrees_scan (zeroed by sort_cells)
  int32_t* flags;      // [B]     probe: some tool receiver kept a non-tool sender (graph.py:135)
  int32_t* n_tools;    // [B]
  int32_t* tools;      // [B*N]
  float* sx; float* sy; float* sz;   // [B*N] particles permuted into sorted order (SoA)
  int32_t* scell;      // [B*N] cell id of each sorted slot
  int32_t* sidx;       // [B*N] original particle id of each sorted slot
  uint8_t* sflag;      // [B*N] bit0 valid, bit1 tool
  int32_t* cell_start; // [B][GRID_MAX_CELLS + 1] first sorted slot of every cell (entries past the graph's cell count = N)
  int32_t* grid_dims;  // [B][4] cells along the two grid axes (a = slow, b = fast), rings to search, bits of (0.998 * min cell width)^2
};

static size_t graph_ws_carve(void* base, int B, int N, int topk, GraphWs* ws) {
  Carver c(base);
  size_t rows = (size_t)B * N;
  size_t nblk = (rows + SCAN_BLOCK - 1) / SCAN_BLOCK;
  GraphWs w;
  w.cand = c.take<int32_t>(rows * topk);
  w.cnt = c.take<int32_t>(rows);
  w.deg = c.take<int32_t>(rows);
  w.lpre = c.take<int32_t>(rows);
  w.blk = c.take<int32_t>(nblk + 1);
  w.total = c.take<int32_t>(1);
  w.ticket = c.take<int32_t>(1);
  w.flags = c.take<int32_t>(B);
  w.n_tools = c.take<int32_t>(B);
  w.tools = c.take<int32_t>(rows);
  w.sx = c.take<float>(rows); w.sy = c.take<float>(rows); w.sz = c.take<float>(rows);
  w.scell = c.take<int32_t>(rows);
  w.sidx = c.take<int32_t>(rows);
  w.sflag = c.take<uint8_t>(rows);
  w.cell_start = c.take<int32_t>((size_t)B * (GRID_MAX_CELLS + 1));
  w.grid_dims = c.take<int32_t>((size_t)B * 4);
  if (ws) *ws = w;
  return align_up(c.off, 256);
}

// ------------------------------------------------------------------------------------ G0
// Ascending list of the graph's tool particles (ordered compaction), by the graph's CTA of T threads.
template <int T>
__device__ __forceinline__ void tool_list_block(const uint8_t* __restrict__ tm, int N, int32_t* __restrict__ tools,
                                                int32_t* __restrict__ n_tools_b, int* warp_tot /*[32] shared*/, int* base_s /*shared*/) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) *base_s = 0;
  __syncthreads();
  for (int j0 = 0; j0 < N; j0 += T) {
    const int j = j0 + tid;
    const bool f = (j < N) && tm[j];
    const unsigned m = __ballot_sync(FULL, f);
    if (lane == 0) warp_tot[warp] = __popc(m);
    __syncthreads();
    int off = *base_s;
    for (int w = 0; w < warp; ++w) off += warp_tot[w];
    if (f) tools[off + __popc(m & ((1u << lane) - 1))] = j;
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < T / 32; ++w) t += warp_tot[w];
      *base_s += t;
    }
    __syncthreads();
  }
  if (tid == 0) *n_tools_b = *base_s;
}

// ------------------------------------------------------------------------------------ G0b
// One CTA per graph.  Grid over the two axes of largest extent of the valid particles; cell width >= the radius (with slack
// for the fp32 rounding of the reference's distance and of the cell coordinate itself), so that two particles within the
// radius differ by at most one cell along each axis.  Particles are counting-sorted by cell id (a * nb + b): histogram with
// shared-memory atomics, block scan (= the first slot of every cell), scatter.
__device__ __forceinline__ int cell_coord(float x, float lo, float w, int n) {
  const float u = __fdiv_rn(__fsub_rn(x, lo), w);
  return min(max((int)fminf(fmaxf(u, 0.f), (float)GRID_MAX_AXIS), 0), n - 1);   // NaN -> 0
}

// T threads per graph: 1024, or 256 for the small graphs of the planning configurations (200 particles x thousands of samples:
// a 1024-thread CTA per graph left most lanes idle and its scan dominated -- 0.11 ms of a 0.25 ms build at 200 x 4096).
template <int T>
__global__ void __launch_bounds__(T) sort_cells_kernel(const float* __restrict__ pos, int64_t pos_stride_b,
                                                           const uint8_t* __restrict__ mask, const uint8_t* __restrict__ tool_mask,
                                                           const float* __restrict__ thr2, int N, float* __restrict__ sx,
                                                           float* __restrict__ sy, float* __restrict__ sz, int32_t* __restrict__ scell,
                                                           int32_t* __restrict__ sidx, uint8_t* __restrict__ sflag,
                                                           int32_t* __restrict__ cell_start, int32_t* __restrict__ grid_dims,
                                                           int32_t* __restrict__ tools, int32_t* __restrict__ n_tools,
                                                           int32_t* __restrict__ flags, int32_t* __restrict__ scan_ticket) {
  __shared__ int cnt[GRID_MAX_CELLS + 1];
  __shared__ int wsum[32];
  __shared__ float red[6][32];
  __shared__ float lo_s[2], w_s[2];
  __shared__ int axis_s[2], n_s[2];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_wait();      // the positions come from the previous kernel of the stream (common.cuh: programmatic dependent launch)
  pdl_trigger();
  const float* p = pos + (size_t)b * pos_stride_b;
  const uint8_t* mk = mask + (size_t)b * N;
  const float INF = __int_as_float(0x7f800000);
  float mn[3] = {INF, INF, INF}, mx[3] = {-INF, -INF, -INF};
  for (int j = tid; j < N; j += T) {
    if (!mk[j]) continue;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float v = p[3 * j + a];
      mn[a] = fminf(mn[a], v);
      mx[a] = fmaxf(mx[a], v);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(FULL, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(FULL, mx[a], o));
    }
    if (lane == 0) { red[a][warp] = mn[a]; red[3 + a][warp] = mx[a]; }
  }
  __syncthreads();
  if (tid == 0) {
    float ext[3], lo3[3], amax = 0.f;
    for (int a = 0; a < 3; ++a) {
      float lo = INF, hi = -INF;
      for (int w = 0; w < T / 32; ++w) { lo = fminf(lo, red[a][w]); hi = fmaxf(hi, red[3 + a][w]); }
      const bool ok = lo <= hi && hi < INF && lo > -INF;   // no valid particle, or non-finite coordinates: a single cell on this axis
      ext[a] = ok ? hi - lo : 0.f;
      lo3[a] = ok ? lo : 0.f;
      if (ok) amax = fmaxf(amax, fmaxf(fabsf(lo), fabsf(hi)));
    }
    const int a0 = (ext[0] >= ext[1] && ext[0] >= ext[2]) ? 0 : (ext[1] >= ext[2] ? 1 : 2);
    const int r1 = (a0 + 1) % 3, r2 = (a0 + 2) % 3;
    const int a1 = ext[r1] >= ext[r2] ? r1 : r2;
    // every sender with (dis - thr^2) < 0 has |x_i - x_j| <= sqrt(thr^2) up to fp32 rounding: widen by 0.1 % + coordinate ulps
    const float hw = sqrtf(fmaxf(thr2[b], 0.f)) * 1.001f + amax * 1e-6f + 1e-30f;
    const int ax[2] = {a0, a1};
    // cells of hw / GRID_SUBDIV: two particles within the radius differ by at most `rings` cells along each axis, where
    // rings * w >= hw (GRID_SUBDIV unless the 64-cell limit widened the cells)
    int rings = 1;
    float wmin = __int_as_float(0x7f800000);
    for (int g = 0; g < 2; ++g) {
      const float e = ext[ax[g]];
      const float w = fmaxf(hw * (1.f / GRID_SUBDIV), e * (1.0001f / GRID_MAX_AXIS));
      const float cells = floorf(e / w) + 1.f;                       // <= GRID_MAX_AXIS by construction
      axis_s[g] = ax[g]; lo_s[g] = lo3[ax[g]]; w_s[g] = w;
      n_s[g] = !(cells >= 1.f) ? 1 : (cells > (float)GRID_MAX_AXIS ? GRID_MAX_AXIS : (int)cells);
      int rg = 1;
      while (rg < GRID_SUBDIV && (float)rg * w < hw) ++rg;          // smallest count with rg * w >= hw (w >= hw / GRID_SUBDIV)
      rings = max(rings, rg);
      wmin = fminf(wmin, w);
    }
    grid_dims[4 * b] = n_s[0];
    grid_dims[4 * b + 1] = n_s[1];
    grid_dims[4 * b + 2] = rings;
    // a sender outside the rings 0..m is >= m cell widths away along a grid axis; 0.2 % slack covers the fp32 rounding of the
    // cell coordinate (same kind of slack as hw above) and of the distance itself
    grid_dims[4 * b + 3] = __float_as_int((wmin * 0.998f) * (wmin * 0.998f));
  }
  __syncthreads();
  const int axa = axis_s[0], axb = axis_s[1], na = n_s[0], nb = n_s[1];
  const float loa = lo_s[0], lob = lo_s[1], wa = w_s[0], wb = w_s[1];
  const int n_cells = na * nb;
  auto cell_of = [&](int j) { return cell_coord(p[3 * j + axa], loa, wa, na) * nb + cell_coord(p[3 * j + axb], lob, wb, nb); };
  // counting sort by cell id.  The order INSIDE a cell is whatever the atomics give: nothing downstream depends on it (knn_rows
  // ranks candidates by (distance, particle id), a strict total order, and emits them sorted by id).
  for (int c = tid; c <= n_cells; c += T) cnt[c] = 0;
  __syncthreads();
  for (int j = tid; j < N; j += T) atomicAdd(&cnt[cell_of(j)], 1);
  __syncthreads();
  // exclusive scan of cnt[0 .. n_cells] (<= 4097 entries): each thread owns PER consecutive entries
  constexpr int PER = (GRID_MAX_CELLS + 1 + T - 1) / T;
  int local[PER], sum = 0;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = tid * PER + i;
    local[i] = c <= n_cells ? cnt[c] : 0;
    sum += local[i];
  }
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = lane < T / 32 ? wsum[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(FULL, w, o);
      if (lane >= o) w += v;
    }
    wsum[lane] = w;
  }
  __syncthreads();
  int run = (warp ? wsum[warp - 1] : 0) + incl - sum;
  int32_t* cs = cell_start + (size_t)b * (GRID_MAX_CELLS + 1);
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = tid * PER + i;
    if (c <= n_cells) { cnt[c] = run; cs[c] = run; }   // cnt becomes the cell's write cursor; cs[n_cells] = N
    run += local[i];
  }
  __syncthreads();
  for (int j = tid; j < N; j += T) {
    const int c = cell_of(j);
    const int s = atomicAdd(&cnt[c], 1);
    const size_t o = (size_t)b * N + s;
    sx[o] = p[3 * j + 0]; sy[o] = p[3 * j + 1]; sz[o] = p[3 * j + 2];
    scell[o] = c;
    sidx[o] = j;
    sflag[o] = (mk[j] ? 1 : 0) | (tool_mask[(size_t)b * N + j] ? 2 : 0);
  }
  if (b == 0 && tid == 0) *scan_ticket = 0;     // degrees_scan's last-block-done counter (same stream, earlier kernel)
  if (tools) {                                   // connect_tools_all
    if (tid == 0) flags[b] = 0;
    __syncthreads();                             // wsum / cnt are free again
    tool_list_block<T>(tool_mask + (size_t)b * N, N, tools + (size_t)b * N, n_tools + b, wsum, &cnt[0]);
  }
}

// ------------------------------------------------------------------------------------ G1
// One THREAD per receiver (a CTA = G1_THREADS consecutive sorted slots, i.e. a few neighbouring cells): the thread walks its
// three slot runs — threads of the same cell read the same addresses (broadcast) — and keeps the k nearest
// in-radius senders in a private sorted list (distance, then sender id) that lives in shared memory, interleaved by thread so
// that list accesses are conflict free.  An insertion only happens when a sender beats the current k-th best, so after the
// first few candidates the loop is pure distance arithmetic.  The list is finally re-sorted by sender id (<= k entries).
// [the warp-per-receiver variant this replaces spent ~1200 issue slots per receiver on shuffles, ballots and a 32-lane
//  bitonic sort: 0.28 ms for 256 k receivers; this one needs ~150]
__global__ void __launch_bounds__(G1_THREADS) knn_rows_kernel(
    const float* __restrict__ sx, const float* __restrict__ sy, const float* __restrict__ sz, const int32_t* __restrict__ scell,
    const int32_t* __restrict__ sidx, const uint8_t* __restrict__ sflag, const int32_t* __restrict__ cell_start,
    const int32_t* __restrict__ grid_dims, const float* __restrict__ thr2, int N, int topk,
    int probe_tools, int smem_cap, int32_t* __restrict__ cand, int32_t* __restrict__ cnt_out, int32_t* __restrict__ flags) {
  extern __shared__ float smem[];
  const int b = blockIdx.y, tid = threadIdx.x;
  const size_t gb = (size_t)b * N;
  pdl_wait();
  pdl_trigger();
  const float t2 = thr2[b];
  const int na = grid_dims[4 * b], nb = grid_dims[4 * b + 1], rings = grid_dims[4 * b + 2];
  const float wq = __int_as_float(grid_dims[4 * b + 3]);
  const int32_t* cs = cell_start + (size_t)b * (GRID_MAX_CELLS + 1);
  const int s_beg = blockIdx.x * G1_ROWS_PER_CTA, s_end = min(N, s_beg + G1_ROWS_PER_CTA);
  // r_lo: a slot at or before everything this CTA's receivers can reach (start of grid row a_first - rings); keeps the offsets small
  const int a_first = scell[gb + s_beg] / nb;
  const int r_lo = cs[max(a_first - rings, 0) * nb];
  (void)smem_cap;
  float* ld = smem;                                                   // [topk][G1_THREADS]
  int32_t* lj = reinterpret_cast<int32_t*>(ld + topk * G1_THREADS);   // [topk][G1_THREADS]  (sender id << 1) | tool bit
  // candidates straight from global memory (read-only and shared by the neighbouring threads, so they are L1 hits); staging the
  // CTA's reachable slot range in shared memory first was measured slower everywhere but on tiny graphs, and 4x slower at 8192
  // particles per graph (the worst-case allocation leaves one CTA per SM): 0.437 -> 0.106 ms for cloth-8192 x 32
  const float* px = sx + gb + r_lo;
  const float* py = sy + gb + r_lo;
  const float* pz = sz + gb + r_lo;
  const int32_t* pj = sidx + gb + r_lo;
  const uint8_t* fl = sflag + gb + r_lo;

  const int slot = s_beg + tid;
  if (slot >= s_end) return;
  const int li = slot - r_lo;
  const int fi = fl[li];
  const int i = pj[li];
  int cnt = 0;
  if (fi & 1) {
    const float xi = px[li], yi = py[li], zi = pz[li];
    const bool tool_i = fi & 2;
    const int ci = scell[gb + slot], ca = ci / nb, cb = ci - ca * nb;
    float kd = __int_as_float(0x7f800000);   // current k-th best (+inf until the list is full)
    int kj = 0x7fffffff;
    // one contiguous run of sorted slots: cells [b0, b1] of grid row a
    auto scan_run = [&](int a, int b0, int b1) {
      const int w_lo = cs[a * nb + b0] - r_lo, w_hi = cs[a * nb + b1 + 1] - r_lo;
      for (int s = w_lo; s < w_hi; ++s) {
        const int fj = fl[s];
        const float dx = __fsub_rn(xi, px[s]), dy = __fsub_rn(yi, py[s]), dz = __fsub_rn(zi, pz[s]);
        const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        const bool ok = (fj & 1) && !(tool_i && (fj & 2)) && (__fsub_rn(d, t2) < 0.f);
        if (!ok) continue;
        const int jc = (pj[s] << 1) | ((fj >> 1) & 1);   // sender id, tool bit in the LSB (order by id is preserved)
        if (!((d < kd) || (d == kd && jc < kj))) continue;
        // sorted insertion (distance, then id); a full list drops its last entry
        int p = cnt < topk ? cnt : topk - 1;
        while (p > 0) {
          const float pd = ld[(p - 1) * G1_THREADS + tid];
          const int pjj = lj[(p - 1) * G1_THREADS + tid];
          if ((pd < d) || (pd == d && pjj < jc)) break;
          ld[p * G1_THREADS + tid] = pd;
          lj[p * G1_THREADS + tid] = pjj;
          --p;
        }
        ld[p * G1_THREADS + tid] = d;
        lj[p * G1_THREADS + tid] = jc;
        if (cnt < topk) ++cnt;
        if (cnt == topk) { kd = ld[(topk - 1) * G1_THREADS + tid]; kj = lj[(topk - 1) * G1_THREADS + tid]; }
      }
    };
    for (int m = 0; m <= rings; ++m) {
      // ring m: the grid rows ca - m and ca + m over cells cb - m .. cb + m, and the two end cells of the rows in between
      for (int a = ca - m; a <= ca + m; ++a) {
        if (a < 0 || a >= na) continue;
        if (a == ca - m || a == ca + m) {
          scan_run(a, max(cb - m, 0), min(cb + m, nb - 1));
        } else {
          if (cb - m >= 0) scan_run(a, cb - m, cb - m);
          if (cb + m < nb) scan_run(a, cb + m, cb + m);
        }
      }
      // every sender outside the rings 0..m is at least m cell widths away: the list is final once its k-th entry is closer
      if (cnt == topk && kd < (float)(m * m) * wq) break;
    }
  }
  // cnt == min(total, topk); re-sort the kept senders by id (insertion sort on <= topk entries)
  int nt = 0;
  for (int t = 0; t < cnt; ++t) {
    const int key = lj[t * G1_THREADS + tid];
    nt += !(key & 1);                            // non-tool candidates (needed for the degree)
    int p = t;
    while (p > 0) {
      const int prev = lj[(p - 1) * G1_THREADS + tid];
      if (prev < key) break;
      lj[p * G1_THREADS + tid] = prev;
      --p;
    }
    lj[p * G1_THREADS + tid] = key;
  }
  const size_t row = gb + i;
  for (int t = 0; t < topk; ++t) cand[row * topk + t] = t < cnt ? (lj[t * G1_THREADS + tid] >> 1) : 0x3fffffff;
  cnt_out[row] = cnt | (nt << 8);
  if (probe_tools && (fi & 2) && cnt > 0) atomicOr(&flags[b], 1);
}

// ------------------------------------------------------------------------------------ G2
__device__ __forceinline__ int row_degree(int packed, bool valid, bool tool_i, int cta, int sem, int flag, int ntools,
                                          bool* keep_base, bool* add_tools) {
  const int cnt = packed & 0xff, nt = packed >> 8;
  if (!cta) { *keep_base = true; *add_tools = false; return cnt; }
  // graph.py:134-144 (batched) / :77-80 (single), see the derivation in DESIGN.md
  *keep_base = !tool_i;
  *add_tools = valid && (sem == AGX_SEM_BATCH ? (flag != 0) : !tool_i);
  return (tool_i ? 0 : nt) + (*add_tools ? ntools : 0);
}

__global__ void __launch_bounds__(SCAN_BLOCK) degrees_scan_kernel(
    const int32_t* __restrict__ cnt, const uint8_t* __restrict__ mask, const uint8_t* __restrict__ tool_mask,
    const int32_t* __restrict__ flags, const int32_t* __restrict__ n_tools, int rows, int N, int cta, int sem,
    int32_t* __restrict__ deg, int32_t* __restrict__ lpre, int32_t* __restrict__ blk, int32_t* __restrict__ ticket,
    int32_t* __restrict__ total, int32_t* __restrict__ row_ptr_end) {
  __shared__ int warp_sum[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r = blockIdx.x * SCAN_BLOCK + tid;
  int d = 0;
  pdl_wait();
  pdl_trigger();
  if (r < rows) {
    const int b = r / N;
    bool kb, at;
    d = row_degree(cnt[r], mask[r] != 0, tool_mask[r] != 0, cta, sem, cta ? flags[b] : 0, cta ? n_tools[b] : 0, &kb, &at);
    deg[r] = d;
  }
  int incl = d;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) warp_sum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = warp_sum[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(FULL, w, o);
      if (lane >= o) w += v;
    }
    warp_sum[lane] = w;  // inclusive over warps
  }
  __syncthreads();
  const int warp_off = warp ? warp_sum[warp - 1] : 0;
  if (r < rows) lpre[r] = warp_off + incl - d;
  if (tid == SCAN_BLOCK - 1) blk[blockIdx.x] = warp_off + incl;
  // ---- the last block to finish turns the block sums into exclusive block offsets (what scan_blocks_kernel used to do)
  __shared__ int is_last;
  __threadfence();
  __syncthreads();
  if (tid == 0) is_last = atomicAdd(ticket, 1) == (int)gridDim.x - 1;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const int nblk = gridDim.x;
  __shared__ int carry_s;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int i0 = 0; i0 < nblk; i0 += SCAN_BLOCK) {
    const int i = i0 + tid;
    const int v = i < nblk ? __ldcg(blk + i) : 0;
    int inc2 = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(FULL, inc2, o);
      if (lane >= o) inc2 += u;
    }
    __syncthreads();
    if (lane == 31) warp_sum[warp] = inc2;
    __syncthreads();
    if (warp == 0) {
      int w = warp_sum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(FULL, w, o);
        if (lane >= o) w += u;
      }
      warp_sum[lane] = w;
    }
    __syncthreads();
    const int carry = carry_s;
    const int excl = carry + (warp ? warp_sum[warp - 1] : 0) + inc2 - v;
    if (i < nblk) blk[i] = excl;
    __syncthreads();
    if (tid == SCAN_BLOCK - 1) carry_s = excl + v;
    __syncthreads();
  }
  if (tid == 0) { *total = carry_s; *row_ptr_end = carry_s; }
}

// ------------------------------------------------------------------------------------ G3
// One thread per receiver: a row has at most topk + n_tools relations, merged in ascending sender order.
__global__ void __launch_bounds__(256) fill_rows_kernel(
    const int32_t* __restrict__ cand, const int32_t* __restrict__ cnt, const int32_t* __restrict__ lpre,
    const int32_t* __restrict__ blk, const int32_t* __restrict__ total, const uint8_t* __restrict__ mask,
    const uint8_t* __restrict__ tool_mask, const int32_t* __restrict__ flags, const int32_t* __restrict__ n_tools,
    const int32_t* __restrict__ tools, int rows, int N, int topk, int cta, int sem, int32_t* __restrict__ row_ptr,
    int32_t* __restrict__ send, int32_t* __restrict__ recv, int64_t cap, int32_t* __restrict__ n_edges,
    int32_t* __restrict__ status) {
  const int r = blockIdx.x * 256 + threadIdx.x;
  pdl_wait();
  pdl_trigger();
  if (r >= rows) return;
  const int b = r / N, n = r - b * N;
  const int start = lpre[r] + blk[r / SCAN_BLOCK];
  row_ptr[r] = start;
  if (n == 0) {
    const int r2 = r + N;
    const int end = (r2 < rows) ? lpre[r2] + blk[r2 / SCAN_BLOCK] : *total;
    n_edges[b] = end - start;
  }
  const int packed = cnt[r];
  const int c = packed & 0xff;
  const bool valid = mask[r] != 0, tool_i = tool_mask[r] != 0;
  const int ntl = cta ? n_tools[b] : 0;
  bool keep_base, add_tools;
  const int deg = row_degree(packed, valid, tool_i, cta, sem, cta ? flags[b] : 0, ntl, &keep_base, &add_tools);
  if (deg == 0) return;
  const int32_t* cd = cand + (size_t)r * topk;
  const int32_t* tl = tools + (size_t)b * N;
  const uint8_t* tm = tool_mask + (size_t)b * N;
  int64_t o = start;
  bool overflow = false;
  int ci = 0, ti = 0;
  const int nt_add = add_tools ? ntl : 0;
  // two-way merge of the (sorted) kept candidates and the (sorted) tool list; the sets are disjoint under cta
  while (true) {
    int cj = 0x7fffffff;
    while (ci < c) {   // next kept candidate
      const int j = cd[ci];
      if (!cta || (keep_base && !tm[j])) { cj = j; break; }
      ++ci;
    }
    const int tj = ti < nt_add ? tl[ti] : 0x7fffffff;
    if (cj == 0x7fffffff && tj == 0x7fffffff) break;
    int j;
    if (cj < tj) { j = cj; ++ci; } else { j = tj; ++ti; }
    if (o < cap) { send[o] = j; recv[o] = r; } else overflow = true;
    ++o;
  }
  if (overflow) atomicOr(status, 1);
}

// ------------------------------------------------------------------------------------ dense one-hot <-> ids
__global__ void __launch_bounds__(256) onehot_to_ids_kernel(const float* __restrict__ R, int64_t rows, int N,
                                                             int32_t* __restrict__ ids) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* p = R + row * N;
  int first = 0x7fffffff;
  for (int j = lane; j < N; j += 32)
    if (p[j] != 0.f) first = min(first, j);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(FULL, first, o));
  if (lane == 0) ids[row] = (first == 0x7fffffff) ? -1 : first;
}

__global__ void __launch_bounds__(256) edges_to_onehot_kernel(const int32_t* __restrict__ row_ptr,
                                                               const int32_t* __restrict__ send, int rows, int N,
                                                               int n_rel, float* __restrict__ Rr, float* __restrict__ Rs) {
  const int r = blockIdx.x * 256 + threadIdx.x;
  if (r >= rows) return;
  const int b = r / N, n = r - b * N;
  const int base = row_ptr[(size_t)b * N];
  for (int e = row_ptr[r]; e < row_ptr[r + 1]; ++e) {
    const int el = e - base;
    if (el < n_rel) {
      Rr[((size_t)b * n_rel + el) * N + n] = 1.f;
      Rs[((size_t)b * n_rel + el) * N + send[e]] = 1.f;
    }
  }
}

// Internal entry shared with the rollout driver (pos may be a strided view of the state history).
int graph_build_impl(const float* pos, int64_t pos_stride_b, const uint8_t* mask, const uint8_t* tool_mask,
                     const float* thr2, int B, int N, int topk, int cta, int sem, int32_t* row_ptr, int32_t* send,
                     int32_t* recv, int64_t cap, int32_t* n_edges, int32_t* status, void* workspace,
                     size_t workspace_bytes, cudaStream_t st) {
  AGX_REQUIRE(B > 0 && N > 0, AGX_ERR_ARG, "graph_build: B=%d N=%d must be positive", B, N);
  AGX_REQUIRE(topk >= 1, AGX_ERR_ARG, "graph_build: topk=%d must be >= 1", topk);
  topk = topk < N ? topk : N;  // graph.py:128
  AGX_REQUIRE(topk <= AGX_MAX_TOPK, AGX_ERR_ARG, "graph_build: min(topk, N)=%d exceeds AGX_MAX_TOPK=%d", topk, AGX_MAX_TOPK);
  AGX_REQUIRE((int64_t)B * N < (1ll << 31) - 1, AGX_ERR_ARG, "graph_build: B*N overflows int32");
  AGX_REQUIRE(sem == AGX_SEM_BATCH || sem == AGX_SEM_SINGLE, AGX_ERR_ARG, "graph_build: bad semantics %d", sem);
  GraphWs ws;
  const size_t need = graph_ws_carve(workspace, B, N, topk, &ws);
  AGX_REQUIRE(workspace && workspace_bytes >= need, AGX_ERR_CAPACITY, "graph_build: workspace %zu < %zu bytes",
              workspace_bytes, need);
  const size_t smem = (size_t)topk * G1_THREADS * 8;   // the per-thread candidate lists of knn_rows
  static thread_local size_t smem_set = 0;
  static thread_local DeviceOnce once;
  if (once.need()) smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    AGX_CUDA_OK(cudaFuncSetAttribute(knn_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  const int rows = B * N;
  const int nblk = (rows + SCAN_BLOCK - 1) / SCAN_BLOCK;
  { ProfScope ps(AGX_KIND_GRAPH_SORT, st);
    if (N <= 512)
      launch_pdl(PDL_SMALL, sort_cells_kernel<256>, B, 256, 0, st, pos, pos_stride_b, mask, tool_mask, thr2, N, ws.sx, ws.sy, ws.sz, ws.scell, ws.sidx,
                                                ws.sflag, ws.cell_start, ws.grid_dims, cta ? ws.tools : nullptr, ws.n_tools, ws.flags, ws.ticket);
    else
      launch_pdl(PDL_SMALL, sort_cells_kernel<1024>, B, 1024, 0, st, pos, pos_stride_b, mask, tool_mask, thr2, N, ws.sx, ws.sy, ws.sz, ws.scell, ws.sidx,
                                                  ws.sflag, ws.cell_start, ws.grid_dims, cta ? ws.tools : nullptr, ws.n_tools, ws.flags, ws.ticket); }
  AGX_LAUNCH_CHECK();
  dim3 g1((N + G1_ROWS_PER_CTA - 1) / G1_ROWS_PER_CTA, B);
  { ProfScope ps(AGX_KIND_GRAPH_KNN, st);
    launch_pdl(PDL_SMALL, knn_rows_kernel, g1, G1_THREADS, smem, st, ws.sx, ws.sy, ws.sz, ws.scell, ws.sidx, ws.sflag, ws.cell_start, ws.grid_dims, thr2, N, topk,
                                                   cta && sem == AGX_SEM_BATCH, N, ws.cand, ws.cnt, ws.flags); }
  AGX_LAUNCH_CHECK();
  { ProfScope ps(AGX_KIND_GRAPH_SCAN, st);
    launch_pdl(PDL_SMALL, degrees_scan_kernel, nblk, SCAN_BLOCK, 0, st, ws.cnt, mask, tool_mask, ws.flags, ws.n_tools, rows, N, cta, sem,
                                                      ws.deg, ws.lpre, ws.blk, ws.ticket, ws.total, row_ptr + rows); }
  AGX_LAUNCH_CHECK();
  { ProfScope ps(AGX_KIND_GRAPH_FILL, st);
    launch_pdl(PDL_SMALL, fill_rows_kernel, (rows + 255) / 256, 256, 0, st, ws.cand, ws.cnt, ws.lpre, ws.blk, ws.total, mask, tool_mask, ws.flags,
                                                      ws.n_tools, ws.tools, rows, N, topk, cta, sem, row_ptr, send, recv,
                                                      cap, n_edges, status); }
  AGX_LAUNCH_CHECK();
  return AGX_OK;
}

}  // namespace agx

extern "C" {

size_t agx_graph_workspace_bytes(int32_t B, int32_t N, int32_t topk) {
  if (B <= 0 || N <= 0 || topk <= 0) return 0;
  topk = topk < N ? topk : N;
  return agx::graph_ws_carve(nullptr, B, N, topk, nullptr);
}

int agx_graph_build(const float* pos, const uint8_t* mask, const uint8_t* tool_mask, const float* thr2, int32_t B,
                    int32_t N, int32_t topk, int32_t connect_tools_all, int32_t semantics, int32_t* row_ptr,
                    int32_t* send, int32_t* recv, int64_t cap, int32_t* n_edges, int32_t* status, void* workspace,
                    size_t workspace_bytes, agx_stream_t stream) {
  AGX_REQUIRE(pos && mask && tool_mask && thr2 && row_ptr && send && recv && n_edges && status, AGX_ERR_ARG,
              "graph_build: null pointer argument");
  return agx::graph_build_impl(pos, (int64_t)N * 3, mask, tool_mask, thr2, B, N, topk, connect_tools_all != 0, semantics,
                               row_ptr, send, recv, cap, n_edges, status, workspace, workspace_bytes,
                               static_cast<cudaStream_t>(stream));
}

int agx_onehot_to_ids(const float* R, int32_t B, int32_t n_rel, int32_t N, int32_t* ids, agx_stream_t stream) {
  AGX_REQUIRE(R && ids && B > 0 && n_rel >= 0 && N > 0, AGX_ERR_ARG, "onehot_to_ids: bad arguments");
  const int64_t rows = (int64_t)B * n_rel;
  if (rows == 0) return AGX_OK;
  { agx::ProfScope ps(AGX_KIND_OTHER, static_cast<cudaStream_t>(stream));
    agx::onehot_to_ids_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(R, rows, N, ids); }
  AGX_LAUNCH_CHECK();
  return AGX_OK;
}

int agx_edges_to_onehot(const int32_t* row_ptr, const int32_t* send, int32_t B, int32_t N, int32_t n_rel, float* Rr,
                        float* Rs, agx_stream_t stream) {
  AGX_REQUIRE(row_ptr && send && Rr && Rs && B > 0 && N > 0 && n_rel >= 0, AGX_ERR_ARG, "edges_to_onehot: bad arguments");
  const int rows = B * N;
  agx::edges_to_onehot_kernel<<<(rows + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(row_ptr, send, rows, N,
                                                                                                 n_rel, Rr, Rs);
  AGX_LAUNCH_CHECK();
  return AGX_OK;
}

}  // extern "C"
