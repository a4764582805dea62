"""Generate tests/golden/rewards.npz with the UNMODIFIED reference functions of src/planning/losses.py (importable as is:
torch + numpy only) — build container only:  python tests/golden/make_golden_rewards.py
running_cost lives in src/planning/plan.py, whose imports (pyflex, GroundingDINO, ...) are absent; its 25 lines are restated
in oracle/planning_oracle.py and exercised against these same inputs in the tests."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/src")


def main():
    from planning.losses import box_loss, chamfer, cloth_penalty, granular_penalty, rope_penalty
    g = torch.Generator().manual_seed(21)
    out = {}
    for name, (B, N, M, By) in {"shared_target": (12, 37, 50, 1), "batched_target": (5, 64, 20, 5), "one_point": (3, 1, 9, 1),
                                "large": (4, 700, 900, 1)}.items():
        x, y = torch.randn(B, N, 3, generator=g), torch.randn(By, M, 3, generator=g) * 1.5
        out[f"chamfer/{name}/x"], out[f"chamfer/{name}/y"] = x.numpy(), y.numpy()
        out[f"chamfer/{name}/out"] = chamfer(x, y).numpy()
    bsz, L, n = 6, 4, 30
    state = torch.randn(bsz, L, n, 3, generator=g)
    state_cur = torch.randn(n, 3, generator=g)
    action = torch.randn(bsz, L, 4, generator=g)
    target = torch.tensor([[-0.5, 0.6], [-0.4, 0.7]])
    out.update({"state": state.numpy(), "state_cur": state_cur.numpy(), "action": action.numpy(), "target": target.numpy()})
    out["box_loss"] = box_loss(state.reshape(bsz * L, n, 3), target).numpy()
    out["rope_penalty"] = rope_penalty(state, action, state_cur).numpy()
    out["cloth_penalty"] = cloth_penalty(state, action, state_cur).numpy()
    out["granular_penalty"] = granular_penalty(state, action, state_cur).numpy()
    np.savez_compressed(os.path.join(HERE, "rewards.npz"), **out)
    print("wrote rewards.npz:", {k: v.shape for k, v in out.items() if "/" not in k})


if __name__ == "__main__":
    main()
