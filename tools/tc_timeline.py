"""Timeline of CTA 0's MMA-issuing thread in tc_edge_encoder_kernel (debug library with -DAGX_TC_TIMELINE).

    python tools/tc_timeline.py build      # here (no GPU): builds adaptigraph_b200/libagx_timeline.so
    AGX_LIB=adaptigraph_b200/libagx_timeline.so python tools/tc_timeline.py run   # on the GPU box
Slots: 1 layer start, 2 weights ready, 3 part-A accumulator free, 4 first A chunk ready, 5 last A chunk ready,
7 part-A issued+committed, 6 part-B accumulator free, 8 part-B issued+committed.
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
DBG = os.path.join(ROOT, "adaptigraph_b200", "libagx_timeline.so")

if sys.argv[1:] == ["build"]:
    from adaptigraph_b200 import build
    print(build.build(force=True, defines=["AGX_TC_TIMELINE"], out=DBG))
    sys.exit(0)

import torch  # noqa: E402
import adaptigraph_b200 as agx  # noqa: E402
from adaptigraph_b200 import _lib as L, synthetic as syn  # noqa: E402

which = sys.argv[2] if len(sys.argv) > 2 else "edge"
L.lib.agx_debug_set_timeline.argtypes = [C.c_void_p]
w = syn.make_workload("cloth", 2000, 128, seed=1238).to("cuda")
m = agx.DynamicsPredictor(*syn.configs("cloth", 3), "cuda").cuda().eval()
el = agx.build_edges(w.state[:, -1], w.adj_thresh, w.state_mask, w.eef_mask, w.topk, w.connect_tools_all).check()
with torch.no_grad():
    m(**w.graph_dict(), edges=el)
    torch.cuda.synchronize()
    buf = torch.zeros(4096, dtype=torch.int64, device="cuda")
    assert L.lib.agx_debug_set_timeline(C.c_void_p(buf.data_ptr())) == 0
    m(**w.graph_dict(), edges=el)
    torch.cuda.synchronize()
    L.lib.agx_debug_set_timeline(None)
for name, base in (("edge_encoder", 0), ("node_encoder", 1024), ("node_update/head (last launched)", 2048)):
    st = [(s >> 48, s & 0xffffffffffff) for s in buf[base:base + 1024].cpu().tolist() if s]
    print("==", name, "stamps", len(st))
    if not st:
        continue
    t0 = prev = st[0][1]
    line = []
    for slot, t in st[:260]:
        if slot == 1 and line:
            print(" ".join(line)); line = []
        line.append(f"{slot}:{t - t0}(+{t - prev})")
        prev = t
    print(" ".join(line))
