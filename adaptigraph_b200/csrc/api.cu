// Version, error reporting and small host utilities of the C ABI.
#include <stdarg.h>

#include <vector>

#include <stdlib.h>
#include "common.cuh"

namespace agx {

char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int set_err(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

int64_t& launch_counter() {
  static thread_local int64_t n = 0;
  return n;
}

int pdl_mask() {
  static const int mask = [] { const char* e = getenv("AGX_PDL"); return e ? atoi(e) : 3; }();
  return mask;
}

static thread_local int g_sm_share = 0;   // > 0: persistent grids are sized for this many SMs (two half-batches side by side)
void set_sm_share(int n) { g_sm_share = n; }

int num_sms() {
  static thread_local int cached = 0;
  if (cached == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || cached <= 0)
      cached = 148;
    // experiment knob (tools only): cap the persistent grids, e.g. to measure a half batch on half the SMs
    if (const char* e = getenv("AGX_SM_CAP")) { const int v = atoi(e); if (v > 0 && v < cached) cached = v; }
  }
  return g_sm_share > 0 && g_sm_share < cached ? g_sm_share : cached;
}

struct ProfRec { int kind; cudaEvent_t e0, e1; };
struct ProfState {
  bool on = false;
  std::vector<ProfRec> recs;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t get() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
  }
};
static ProfState& prof() {
  static thread_local ProfState p;
  return p;
}

ProfScope::ProfScope(int kind_, cudaStream_t st_) : kind(kind_), st(st_) {
  if (!prof().on) return;
  e0 = prof().get();
  if (e0) cudaEventRecord(e0, st);
}
ProfScope::~ProfScope() {
  if (!e0) return;
  cudaEvent_t e1 = prof().get();
  if (e1) cudaEventRecord(e1, st);
  prof().recs.push_back({kind, e0, e1});
}

PackedLayout packed_layout() {
  PackedLayout L;
  size_t o = 0;
  auto mat = [&](int kpad) { size_t r = o; o += (size_t)kpad * FP; return r; };
  auto vec = [&](int n) { size_t r = o; o += n; return r; };
  L.penc0_w = mat(D_NODE_IN); L.penc0_b = vec(FP);
  L.penc2_w = mat(FP);        L.penc2_b = vec(FP);
  L.penc4_w = mat(FP);        L.penc4_b = vec(FP);
  L.renc0_w = mat(D_REL_IN);  L.renc0_b = vec(FP);
  L.renc2_w = mat(FP);        L.renc2_b = vec(FP);
  L.renc4_w = mat(FP);        L.renc4_b = vec(FP);
  L.rp_rel_w = mat(FP);       L.rp_b = vec(FP);
  L.rp_recv_w = mat(FP);      L.rp_send_w = mat(FP);
  L.pp_enc_w = mat(FP);       L.pp_b = vec(FP);
  L.pp_agg_w = mat(FP);
  L.pred0_w = mat(FP);        L.pred0_b = vec(FP);
  L.pred1_w = mat(FP);        L.pred1_b = vec(FP);
  L.pred2_w = vec(3 * FP);    L.pred2_b = vec(4);
  L.total = o;
  return L;
}

}  // namespace agx

extern "C" {

int agx_version(void) { return AGX_VERSION; }

const char* agx_last_error(void) { return agx::err_buf(); }

int64_t agx_launch_count(void) { return agx::launch_counter(); }

int agx_profile_enable(int32_t on) {
  agx::prof().on = on != 0;
  return AGX_OK;
}

int agx_profile_read(double* ms, int64_t* count) {
  AGX_REQUIRE(ms && count, AGX_ERR_ARG, "profile_read: null pointer argument");
  auto& p = agx::prof();
  for (auto& r : p.recs) {
    if (r.e0 && r.e1) {
      AGX_CUDA_OK(cudaEventSynchronize(r.e1));
      float t = 0.f;
      AGX_CUDA_OK(cudaEventElapsedTime(&t, r.e0, r.e1));
      if (r.kind >= 0 && r.kind < AGX_NUM_KINDS) { ms[r.kind] += t; count[r.kind] += 1; }
    }
    if (r.e0) p.pool.push_back(r.e0);
    if (r.e1) p.pool.push_back(r.e1);
  }
  p.recs.clear();
  return AGX_OK;
}

const char* agx_kind_name(int32_t kind) {
  static const char* names[AGX_NUM_KINDS] = {"graph_tool_list", "graph_knn_rows", "graph_scan", "graph_fill_rows",
                                             "node_encoder", "edge_encoder", "edge_aggregate", "node_update",
                                             "node_update_head", "rollout_advance", "other", "graph_sort_cells"};
  return (kind >= 0 && kind < AGX_NUM_KINDS) ? names[kind] : "?";
}

}  // extern "C"
