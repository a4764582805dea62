"""BASELINE configs[2] and [4]: granular 1k x 64 x 5-step rollout, and the mixed rope / granular / cloth sweep over
256..8192 particles (B = 262144 / n_p graphs per GPU, pstep 3, 10-step rollout with re-graphing).  One JSON line per case."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import adaptigraph_b200 as agx  # noqa: E402
from adaptigraph_b200 import synthetic as syn  # noqa: E402


def run(material, n_p, B, T, pstep=3, iters=3):
    torch.manual_seed(0)
    m = agx.DynamicsPredictor(*syn.configs(material, pstep), "cuda").cuda().eval()
    w = syn.make_workload(material, n_p, B, seed=1240).to("cuda")
    thr, topk, cta, n_s = syn.MATERIALS[material]
    max_nR = (n_p + n_s) * (topk + n_s) + 64
    f = lambda: m.rollout(w.state, w.attrs, w.action, w.p_instance, w.physics_param, w.state_mask, w.eef_mask, thr, topk, cta, T, max_nR, check=False)  # noqa: E731
    out = f(); f()
    torch.cuda.synchronize()
    assert int(out["n_edges"].max()) <= max_nR
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(json.dumps({"material": material, "n_p": n_p, "graphs": B, "rollout_steps": T, "pstep": pstep, "ms_per_rollout": round(ms, 3),
                      "particle_steps_per_s": round(B * n_p * T / (ms * 1e-3)), "relations_per_graph": float(out["n_edges"].float().mean())}), flush=True)
    del m, w, out
    torch.cuda.empty_cache()


if __name__ == "__main__":
    run("granular", 1000, 64, 5)                       # configs[2]
    for material in ("rope", "granular", "cloth"):     # configs[4]
        for n_p in (256, 512, 1024, 2048, 4096, 8192):
            run(material, n_p, 262144 // n_p, 10)
