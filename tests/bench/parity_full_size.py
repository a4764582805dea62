"""Parity at BASELINE configs[3] size beyond the bench's graph 0: cloth 2000 particles, 10-step rollout with re-graphing, several graphs,
engine (tensor-core and fp32 arithmetic) against the oracle's dense CPU restatement: per-step relation counts must be identical,
positions within the north-star tolerance (1e-4 RMSE)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import adaptigraph_b200 as agx  # noqa: E402
from adaptigraph_b200 import synthetic as syn  # noqa: E402
from oracle import dynamics_oracle as orc  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 6
T, K = 10, 3
torch.manual_seed(0)
torch.set_num_threads(os.cpu_count())
m = agx.DynamicsPredictor(*syn.configs("cloth", K), "cuda").cuda().eval()
params = {k: v.detach().cpu() for k, v in m.state_dict().items()}
w = syn.make_workload("cloth", 2000, B, seed=4242)
t0 = time.perf_counter()
refs, edges = [], []
for b in range(B):                                   # one graph at a time: the dense one-hots of a graph are 14k x 2002 floats each
    wb = w.take(slice(b, b + 1))
    r, e = orc.rollout_dense(params, K, wb.state, wb.attrs, wb.p_instance, wb.action, wb.physics_param, wb.state_mask, wb.eef_mask,
                             w.adj_thresh, w.topk, w.connect_tools_all, T)
    refs.append(r)
    edges.append(torch.stack([(Rr.sum(-1) > 0).sum(1) for Rr, _ in e], 0))
ref = torch.cat(refs, 0)
n_ref = torch.cat(edges, 1)
cpu_s = time.perf_counter() - t0
wd = w.to("cuda")
out = {}
for prec in ("tc", "fp32"):
    m.set_precision(prec)
    o = m.rollout(wd.state, wd.attrs, wd.action, wd.p_instance, wd.physics_param, wd.state_mask, wd.eef_mask, w.adj_thresh, w.topk,
                  w.connect_tools_all, T, max_nR=16384)
    err = o["state_seqs"].cpu() - ref
    out[prec] = {"relation_counts_identical": bool(torch.equal(o["n_edges"].cpu().long(), n_ref)),
                 "rmse": float(err.pow(2).mean().sqrt()), "max_abs": float(err.abs().max()),
                 "rmse_per_step": [float(x) for x in err.pow(2).mean(dim=(0, 2, 3)).sqrt()]}
print(json.dumps({"workload": f"cloth 2000 particles x {B} graphs, {T}-step rollout, pstep {K}", "cpu_seconds": round(cpu_s, 1), **out}))
