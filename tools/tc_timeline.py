"""Timeline of CTA 0's MMA-issuing thread in tc_edge_encoder_kernel (debug library with -DAGX_TC_TIMELINE).

    python tools/tc_timeline.py build edge|node|update|head    # here (no GPU): builds adaptigraph_b200/libagx_timeline_<k>.so
    AGX_LIB=adaptigraph_b200/libagx_timeline_edge.so python tools/tc_timeline.py run   # on the GPU box
Columns: the two MMA threads, the weight loader, lane 0 of epilogue warps 0 and 4 (slot 0's two column halves).
MMA: L layer start, w weights ready, eA part-A accumulator free, in A ready, A! part A issued, eB part-B accumulator free, B! issued.
Loader: lw waiting for a free buffer, le got it (copy issued).  Epilogue: T tile start, pr producer done, l layer start,
fA part A full, rA read+released, mA part-A math done, fB part B full, sB stores issued, sg signalled, o/O last layer begin/end.
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
DBG = os.path.join(ROOT, "adaptigraph_b200", "libagx_timeline.so")

KIDS = {"edge": 4, "node": 6, "update": 3, "head": 13}
if sys.argv[1:2] == ["build"]:
    from adaptigraph_b200 import build
    which = sys.argv[2] if len(sys.argv) > 2 else "edge"
    print(build.build(force=True, defines=[f"AGX_TC_TIMELINE={KIDS[which]}"], out=DBG.replace(".so", f"_{which}.so")))
    sys.exit(0)

import torch  # noqa: E402
import adaptigraph_b200 as agx  # noqa: E402
from adaptigraph_b200 import _lib as L, synthetic as syn  # noqa: E402

L.lib.agx_debug_set_timeline.argtypes = [C.c_void_p]
w = syn.make_workload("cloth", 2000, 128, seed=1238).to("cuda")
m = agx.DynamicsPredictor(*syn.configs("cloth", 3), "cuda").cuda().eval()
el = agx.build_edges(w.state[:, -1], w.adj_thresh, w.state_mask, w.eef_mask, w.topk, w.connect_tools_all).check()
with torch.no_grad():
    m(**w.graph_dict(), edges=el)
    torch.cuda.synchronize()
    buf = torch.zeros(4000, dtype=torch.int64, device="cuda")
    assert L.lib.agx_debug_set_timeline(C.c_void_p(buf.data_ptr())) == 0
    m(**w.graph_dict(), edges=el)
    torch.cuda.synchronize()
    L.lib.agx_debug_set_timeline(None)
names = {1: "L", 2: "w", 3: "eA", 4: "in", 7: "A!", 6: "eB", 8: "B!", 10: "lw", 11: "le", 20: "l", 21: "fA", 22: "rA", 23: "mA",
         24: "fB", 25: "sB", 26: "sg", 30: "T", 31: "pr", 32: "o", 33: "O", 34: "ofA", 35: "orA", 36: "ocA", 37: "ofB", 38: "orB",
         40: "r", 41: "rfA", 43: "rmA", 44: "rfB", 45: "rsB"}
ev = []
t0 = None
for region, rname in enumerate(("mma0", "mma1", "load", "epi0", "epi4")):
    for s in buf[region * 800:(region + 1) * 800].cpu().tolist():
        if s:
            ev.append((s & 0xffffffffffff, rname, names.get(s >> 48, str(s >> 48))))
ev.sort()
t0 = ev[0][0]
skip = [e for e in ev if e[0] - t0 > int(os.environ.get("TL_SKIP", "60000"))]      # steady state: skip the first rounds
cols = ("mma0", "mma1", "load", "epi0", "epi4")
print("clk      " + "".join(f"{c:>8s}" for c in cols))
for t, r, n in skip[:700]:
    print(f"{t - t0:8d} " + "".join(f"{(n if c == r else ''):>8s}" for c in cols))
