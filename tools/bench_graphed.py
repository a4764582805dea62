"""Eager vs CUDA-graph-replayed rollout at small batches (launch-bound regime): cloth-2k, 10 steps, graphs per GPU in argv."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import adaptigraph_b200 as agx  # noqa: E402
from adaptigraph_b200 import synthetic as syn  # noqa: E402

for B in [int(a) for a in sys.argv[1:]] or [8, 16, 32]:
    for material, n_p, T in (("cloth", 2000, 10), ("rope", 100, 10)):
        torch.manual_seed(0)
        m = agx.DynamicsPredictor(*syn.configs(material, 3), "cuda").cuda().eval()
        w = syn.make_workload(material, n_p, B, seed=3).to("cuda")
        args = (w.state, w.attrs, w.action, w.p_instance, w.physics_param, w.state_mask, w.eef_mask, w.adj_thresh, w.topk, w.connect_tools_all)
        max_nR = (n_p + 8) * 12
        eager = lambda: m.rollout(*args, T, max_nR, check=False)  # noqa: E731
        graphed = agx.GraphedRollout(m, *args, T, max_nR)

        def timeit(f, n=20):
            for _ in range(3):
                f()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n):
                f()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / n
        te, tg = timeit(eager), timeit(lambda: graphed(state=w.state))
        print(json.dumps({"material": material, "n_p": n_p, "graphs": B, "rollout_steps": T, "eager_ms": round(te, 3), "graphed_ms": round(tg, 3),
                          "graphed_particle_steps_per_s": round(B * n_p * T / (tg * 1e-3))}), flush=True)
