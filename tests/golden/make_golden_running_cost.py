"""Golden vectors for the planner's reward tail: runs the reference's OWN `running_cost` (src/planning/plan.py:27-59) with the
reference's own criteria and penalties (src/planning/losses.py), bound exactly as plan.py:146-175 binds them.

plan.py cannot be imported here (its module-level imports pull in pyflex, open3d, GroundingDINO, the robot drivers ...), so the
function's source is taken from the file UNMODIFIED — the `def running_cost` node of its syntax tree — and executed in a namespace
that holds `torch` only.  Build container only:  python tests/golden/make_golden_running_cost.py
"""
import ast
import contextlib
import io
import os
import sys
from functools import partial

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src"
sys.path.insert(0, REF)


def reference_running_cost():
    src = open(os.path.join(REF, "planning", "plan.py")).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "running_cost")
    ns = {"torch": torch}
    exec(compile(ast.Module(body=[node], type_ignores=[]), "plan.py", "exec"), ns)   # noqa: S102  (the reference's own function)
    return ns["running_cost"]


def main():
    from planning.losses import box_loss, chamfer, cloth_penalty, granular_penalty, rope_penalty
    running_cost = reference_running_cost()
    g = torch.Generator().manual_seed(33)
    bsz, L, n, M = 7, 3, 40, 55
    out = {}
    state_cur = torch.rand(n, 3, generator=g) * 4 - 2
    state = state_cur[None, None] + 0.3 * torch.randn(bsz, L, n, 3, generator=g)
    action = torch.cat([torch.rand(bsz, L, 2, generator=g) * 6 - 3, torch.rand(bsz, L, 1, generator=g) * 6.28,
                        torch.rand(bsz, L, 1, generator=g)], -1)
    target_pts = torch.rand(M, 3, generator=g) * 4 - 2
    target_box = torch.tensor([[-0.8, 1.1], [-0.6, 0.9]])
    bbox = np.array([[-2.5, 2.5], [-2.2, 2.4]])                      # plan.py:170-174 keeps it a numpy array
    out.update(state=state.numpy(), action=action.numpy(), state_cur=state_cur.numpy(), target_pts=target_pts.numpy(),
               target_box=target_box.numpy(), bbox=bbox.astype(np.float32))
    for ename, crit in (("chamfer", partial(chamfer, y=target_pts[None])), ("box", partial(box_loss, target=target_box))):
        for pname, pen in (("rope", rope_penalty), ("cloth", cloth_penalty), ("granular", granular_penalty)):
            for ratio in (10.0, 4.0):
                with contextlib.redirect_stdout(io.StringIO()):
                    r = running_cost(state, action, state_cur, error_func=crit, penalty_func=partial(pen, sim_real_ratio=ratio), bbox=bbox)
                out[f"reward/{ename}/{pname}/{ratio:g}"] = r["reward_seqs"].numpy()
    np.savez_compressed(os.path.join(HERE, "running_cost.npz"), **out)
    print("wrote running_cost.npz:", len(out), "arrays;", {k: v for k, v in list(out.items())[-1:]})


if __name__ == "__main__":
    main()
