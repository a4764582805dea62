"""Golden vectors for the MPC drivers: runs the UNMODIFIED reference `dynamics` / `dynamics_masked`
(src/planning/forward_dynamics.py) on CPU with a stand-in ppm_optimizer.  Build container only."""
import contextlib
import io
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402


def main():
    from adaptigraph_b200 import synthetic as syn
    DP = import_reference()[0]
    from planning.forward_dynamics import dynamics, dynamics_masked
    torch.set_num_threads(4)
    out = {}
    for name, material, pusher, gripper in [("rope1pt", "rope", [[0.0, 0.0, 0.12]], False),
                                            ("granular5pt", "granular", [[0, 0, 0]] + [[0, v, 0] for v in (0.05, -0.05, 0.025, -0.025)], False),
                                            ("cloth_gripper", "cloth", [[0.0, 0.0, 0.12]], True)]:
        thr, topk, cta, _ = syn.MATERIALS[material]
        mc, matc, dc = syn.configs(material, 3)
        torch.manual_seed(0)
        model = DP(mc, matc, dc, "cpu").eval()
        w = syn.make_workload(material, 40, 1, seed=61, n_s=len(pusher))
        state = w.state[0, -1, :40].clone()
        ppm = types.SimpleNamespace(
            task_config=dict(max_n=1, max_nR=1200, n_his=4, sim_real_ratio=10.0, push_length=0.1, pusher_points=pusher, gripper_enable=gripper,
                             topk=topk, connect_tools_all=cta),
            eef_num=len(pusher), material=material, material_dims={material: 1}, material_indices={material: 0},
            physics_param={material: torch.tensor([0.4])}, adj_thresh=thr)
        g = torch.Generator().manual_seed(62)
        bsz, n_look = 5, 2
        action = torch.zeros(bsz, n_look, 4)
        c = state.mean(0)
        action[:, :, 0] = c[0] + 0.3 * torch.randn(bsz, n_look, generator=g)
        action[:, :, 1] = c[2] + 0.3 * torch.randn(bsz, n_look, generator=g)
        action[:, :, 2] = 6.28 * torch.rand(bsz, n_look, generator=g)
        action[:, :, 3] = torch.randint(1, 5, (bsz, n_look), generator=g).float() + 0.5
        with contextlib.redirect_stdout(io.StringIO()):
            res = dynamics(state, action, model, "cpu", ppm)
        out[f"{name}/state"] = state.numpy(); out[f"{name}/action"] = action.numpy()
        out[f"{name}/state_seqs"] = res["state_seqs"].numpy(); out[f"{name}/action_seqs"] = res["action_seqs"].numpy()
        out[f"{name}/pusher"] = np.asarray(pusher, np.float32); out[f"{name}/gripper"] = np.asarray(gripper); out[f"{name}/material"] = np.asarray(material)
        # masked variant: per-sample states with some particles masked out
        st = state[None].repeat(bsz, 1, 1) + 0.01 * torch.randn(bsz, 40, 3, generator=g)
        mask = torch.ones(bsz, 40, dtype=torch.bool)
        mask[1, 30:] = False
        mask[3, 35:] = False
        with contextlib.redirect_stdout(io.StringIO()):
            res = dynamics_masked(st, mask, action[:, 0], model, "cpu", ppm)
        out[f"{name}/m_state"] = st.numpy(); out[f"{name}/m_mask"] = mask.numpy()
        out[f"{name}/m_state_seqs"] = res["state_seqs"].numpy(); out[f"{name}/m_action_seqs"] = res["action_seqs"].numpy()
        print(name, "ok", float(res["state_seqs"].abs().mean()))
    np.savez_compressed(os.path.join(HERE, "planning_dynamics.npz"), **out)


if __name__ == "__main__":
    main()
