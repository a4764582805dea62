"""Per-kernel table of the device-resident rollout at the reference's planning sizes (config/planning/*.yaml:31-42: max_nobj 200,
n_sample 20000 in chunks of 500): 200 particles per graph, thousands of samples per call."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import adaptigraph_b200 as agx  # noqa: E402
from adaptigraph_b200 import ops, synthetic as syn  # noqa: E402

for material, n_p, B, T in [("rope", 200, 500, 5), ("rope", 200, 4000, 5), ("granular", 200, 4000, 5), ("cloth", 200, 4000, 5), ("rope", 200, 20000, 5)]:
    torch.manual_seed(0)
    m = agx.DynamicsPredictor(*syn.configs(material, 3), "cuda").cuda().eval()
    w = syn.make_workload(material, n_p, B, seed=3).to("cuda")
    thr, topk, cta, n_s = syn.MATERIALS[material]
    max_nR = 2000 if material == "rope" else 4000   # the synthetic granular / cloth graphs are denser than the planning configs assume
    f = lambda: m.rollout(w.state, w.attrs, w.action, w.p_instance, w.physics_param, w.state_mask, w.eef_mask, thr, topk, cta, T, max_nR, check=False)  # noqa: E731
    out = f(); f()
    torch.cuda.synchronize()
    assert int(out["n_edges"].max()) <= max_nR
    ops.profile_read(); ops.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        f()
    e1.record()
    torch.cuda.synchronize()
    prof = ops.profile_read(); ops.profile_enable(False)
    ms = e0.elapsed_time(e1) / 3
    print(json.dumps({"material": material, "n_p": n_p, "samples": B, "rollout_steps": T, "ms_per_rollout": round(ms, 3),
                      "particle_steps_per_s": round(B * n_p * T / (ms * 1e-3)), "relations_per_graph": float(out["n_edges"].float().mean()),
                      "kernel_ms": {k: round(v[0] / v[1], 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}}), flush=True)
    del m, w, out
    torch.cuda.empty_cache()
