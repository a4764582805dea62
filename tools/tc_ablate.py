"""Where the relation encoder's time goes: one thing removed per build (results are garbage, only the kernel time is read).

    python tools/tc_ablate.py build      # here (no GPU): variants/libagx_abl_<what>.so
    python tools/tc_ablate.py run        # on the GPU box: per-kernel milliseconds of one forward on cloth-2k x 128 per variant
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VARIANTS = {"full": [], "no_tmem_ld": ["AGX_ABLATE_TMEM_LD"], "no_tmem_st": ["AGX_ABLATE_TMEM_ST"], "no_mma": ["AGX_ABLATE_MMA"],
            "no_c16_store": ["AGX_ABLATE_C16_STORE"], "no_mma_no_ld": ["AGX_ABLATE_MMA", "AGX_ABLATE_TMEM_LD"]}
path = lambda n: os.path.join(ROOT, "variants", f"libagx_abl_{n}.so")  # noqa: E731

if sys.argv[1:2] == ["build"]:
    from adaptigraph_b200 import build
    os.makedirs(os.path.join(ROOT, "variants"), exist_ok=True)
    for n, d in VARIANTS.items():
        print(build.build(force=True, defines=d, out=path(n)))
    sys.exit(0)

if sys.argv[1:2] == ["run"]:
    for n in VARIANTS:
        r = subprocess.run([sys.executable, __file__, "one"], env=dict(os.environ, AGX_LIB=path(n)), capture_output=True, text=True, timeout=300)
        print(f"{n:14s} {r.stdout.strip() or r.stderr.strip()[-300:]}", flush=True)
    sys.exit(0)

import torch  # noqa: E402
import adaptigraph_b200 as agx  # noqa: E402
from adaptigraph_b200 import ops, synthetic as syn  # noqa: E402

w = syn.make_workload("cloth", 2000, 128, seed=1238).to("cuda")
m = agx.DynamicsPredictor(*syn.configs("cloth", 3), "cuda").cuda().eval()
el = agx.build_edges(w.state[:, -1], w.adj_thresh, w.state_mask, w.eef_mask, w.topk, w.connect_tools_all).check()
with torch.no_grad():
    for _ in range(3):
        m(**w.graph_dict(), edges=el)
    torch.cuda.synchronize()
    ops.profile_read()
    ops.profile_enable(True)
    for _ in range(5):
        m(**w.graph_dict(), edges=el)
    prof = ops.profile_read()
    ops.profile_enable(False)
print("  ".join(f"{k} {ms / max(c, 1):.4f}" for k, (ms, c) in sorted(prof.items()) if k in ("edge_encoder", "node_encoder", "node_update", "node_update_head", "edge_aggregate")))
