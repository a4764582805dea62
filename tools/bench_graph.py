"""Per-kernel times of the graph builder alone over graph sizes from the reference's planning configs (200 particles x thousands of
samples, config/planning/*.yaml:31-42) to 8192 particles per graph."""
import json, sys, torch
sys.path.insert(0, '.')
import adaptigraph_b200 as agx
from adaptigraph_b200 import synthetic as syn, ops
for material, n_p, B in [("cloth", 2000, 128), ("cloth", 8192, 32), ("granular", 4096, 64), ("rope", 512, 512), ("rope", 200, 4096),
                         ("granular", 200, 4096), ("cloth", 200, 4096)]:
    w = syn.make_workload(material, n_p, B, seed=5).to("cuda")
    f = lambda: agx.build_edges(w.state[:, -1], w.adj_thresh, w.state_mask, w.eef_mask, w.topk, w.connect_tools_all, max_nR=(n_p + 8) * 28)
    f(); f(); torch.cuda.synchronize()
    ops.profile_read(); ops.profile_enable(True)
    for _ in range(5): el = f()
    torch.cuda.synchronize()
    p = ops.profile_read(); ops.profile_enable(False)
    print(material, n_p, B, {k: round(v[0] / v[1], 4) for k, v in p.items()}, int(el.row_ptr[-1]))
