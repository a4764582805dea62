// Shared host/device helpers for the adaptigraph_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/adaptigraph_b200.h"

namespace agx {

constexpr int FP = AGX_FP;          // padded feature stride (floats)
constexpr int NFEAT = AGX_NFEAT;    // per-node relation-input record (floats)
constexpr int H_FIX = 4;            // history frames the kernels are specialised for
constexpr int D_NODE_IN = 8;        // node encoder input (6) padded to a multiple of 8
constexpr int D_REL_IN = 24;        // relation encoder input (17) padded to a multiple of 8

// ---- error plumbing (thread-local, no exceptions across the ABI)
char* err_buf();
int set_err(int code, const char* fmt, ...);
int64_t& launch_counter();

#define AGX_CUDA_OK(expr)                                                                   \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return agx::set_err(AGX_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
  } while (0)

#define AGX_LAUNCH_CHECK()                                                                  \
  do {                                                                                      \
    agx::launch_counter()++;                                                                \
    cudaError_t _e = cudaGetLastError();                                                    \
    if (_e != cudaSuccess)                                                                  \
      return agx::set_err(AGX_ERR_CUDA, "%s:%d launch: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
  } while (0)

#define AGX_REQUIRE(cond, code, ...)                                                        \
  do {                                                                                      \
    if (!(cond)) return agx::set_err(code, __VA_ARGS__);                                    \
  } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Carves 256-byte aligned sub-buffers out of the caller's workspace.
struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* p) : base(static_cast<char*>(p)) {}
  template <typename T>
  T* take(size_t n) {
    off = align_up(off, 256);
    T* p = reinterpret_cast<T*>(base + off);
    off += n * sizeof(T);
    return p;
  }
};

int num_sms();
// cudaFuncSetAttribute / stream caches are per device: a host thread that switches devices must redo them
struct DeviceOnce {
  int dev = -1;
  bool need() {   // true the first time it is asked on the calling thread's current device
    int d = -1;
    if (cudaGetDevice(&d) != cudaSuccess) return true;
    if (d == dev) return false;
    dev = d;
    return true;
  }
};
void set_sm_share(int n);   // thread-local: size persistent grids for n SMs (0 = all)

// ---- programmatic dependent launch: a kernel of the rollout / forward chain is launched with the stream-serialisation
// attribute, so its launch (and, once the previous grid's CTAs have exited, its prologue: barrier init, tensor-memory allocation,
// shared-memory set-up) no longer waits for the previous grid's completion to be processed; every such kernel executes
// pdl_wait() before its first global-memory access -- the wait returns once ALL earlier grids have completed and their writes
// are visible, so the ordering guarantees are those of a plain stream.  AGX_PDL is a mask (default 3; 0 = plain launches, the
// device-side instructions are then no-ops).
// Measured (profiles/r02H_pdl_ab.txt, cloth-2k x 128 rollout as a CUDA graph): attribute alone +0.5 % (132.0-132.3 M vs
// 131.1-131.7 M particle-steps/s); with an EARLY griddepcontrol.launch_dependents at the top of the chain / aggregate kernels
// -8 % (121.5 M; -7 % at 16 graphs), on the small kernels +-0 -- so no kernel triggers early (AGX_PDL_EARLY_TRIGGER restores it).
int pdl_mask();   // AGX_PDL: bit 0 = the small kernels (graph builder, rollout_advance), bit 1 = the chain / aggregate kernels
constexpr int PDL_SMALL = 1, PDL_BIG = 2;
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() {
#ifdef AGX_PDL_EARLY_TRIGGER
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(int cls, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl_mask() & cls) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// ---- optional per-kernel timing (agx_profile_*): CUDA events recorded on the launch stream
// around every kernel, summed per kernel kind when read.  Off by default.
struct ProfScope {
  int kind;
  cudaStream_t st;
  cudaEvent_t e0 = nullptr;
  ProfScope(int kind, cudaStream_t st);
  ~ProfScope();
};

// ---- packed weight blob (built by agx_pack_weights) -------------------------------------
// Every matrix is stored k-major ("transposed"): Wt[k*FP + n] = W[n][col0 + k], zero padded to
// [Kpad][FP]; biases are padded to FP floats.  Offsets are in floats.
struct PackedLayout {
  size_t penc0_w, penc0_b, penc2_w, penc2_b, penc4_w, penc4_b;
  size_t renc0_w, renc0_b, renc2_w, renc2_b, renc4_w, renc4_b;
  size_t rp_rel_w, rp_b, rp_recv_w, rp_send_w;   // relation_propagator split by operand
  size_t pp_enc_w, pp_b, pp_agg_w;               // particle_propagator split by operand
  size_t pred0_w, pred0_b, pred1_w, pred1_b;
  size_t pred2_w, pred2_b;                       // pred2_w is row-major [3][FP]; pred2_b 4 floats
  size_t total;
};
PackedLayout packed_layout();

}  // namespace agx
