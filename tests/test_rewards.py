"""MPC reward terms: the oracle and — on the GPU — adaptigraph_b200.rewards against the outputs of the
reference's own src/planning/losses.py (tests/golden/rewards.npz, tests/golden/make_golden_rewards.py)."""
from functools import partial

import numpy as np
import pytest
import torch

import agx_helpers as H
from oracle import planning_oracle as po

G = H.load_npz("rewards.npz")
CH = sorted({k.split("/")[1] for k in G if k.startswith("chamfer/")})
TOL = 2e-6


def test_oracle_rope_penalty_matches_reference():
    t = lambda k: torch.from_numpy(G[k])  # noqa: E731
    np.testing.assert_allclose(po.rope_penalty(t("state"), t("action"), t("state_cur")).numpy(), G["rope_penalty"], rtol=0, atol=1e-7)


@pytest.mark.parametrize("name", CH)
def test_oracle_chamfer_matches_reference(name):
    out = po.chamfer(torch.from_numpy(G[f"chamfer/{name}/x"]), torch.from_numpy(G[f"chamfer/{name}/y"]))
    np.testing.assert_allclose(out.numpy(), G[f"chamfer/{name}/out"], rtol=0, atol=1e-7)


@pytest.fixture(scope="module")
def rw():
    import adaptigraph_b200.ops  # noqa: F401  (loads the .so; raises if missing)
    from adaptigraph_b200 import rewards
    assert torch.cuda.is_available()
    return rewards


@pytest.mark.gpu
@pytest.mark.parametrize("name", CH)
def test_gpu_chamfer_matches_reference(rw, name):
    x, y = torch.from_numpy(G[f"chamfer/{name}/x"]).cuda(), torch.from_numpy(G[f"chamfer/{name}/y"]).cuda()
    out = rw.chamfer(x, y)
    assert out.is_cuda and out.shape == (x.shape[0],)
    np.testing.assert_allclose(out.cpu().numpy(), G[f"chamfer/{name}/out"], rtol=TOL, atol=TOL)


@pytest.mark.gpu
def test_gpu_chamfer_rejects_bad_input(rw):
    with pytest.raises(RuntimeError):
        rw.chamfer(torch.zeros(2, 4, 3), torch.zeros(1, 4, 3))                       # CPU tensors: no fallback
    with pytest.raises(RuntimeError):
        rw.rope_penalty(torch.zeros(2, 3, 4, 3), torch.zeros(2, 3, 4), torch.zeros(4, 3))
    with pytest.raises(ValueError):
        rw.chamfer(torch.zeros(2, 4, 3).cuda(), torch.zeros(3, 4, 3).cuda())         # target batch neither 1 nor B
    with pytest.raises(ValueError):
        rw.chamfer(torch.zeros(1, 9000, 3).cuda(), torch.zeros(1, 9000, 3).cuda())   # beyond the shared-memory staging limit


@pytest.mark.gpu
def test_gpu_penalties_and_running_cost_match_reference(rw):
    t = lambda k: torch.from_numpy(G[k]).cuda()  # noqa: E731
    state, action, state_cur, target = t("state"), t("action"), t("state_cur"), t("target")
    bsz, L, n, _ = state.shape
    np.testing.assert_allclose(rw.box_loss(state.reshape(bsz * L, n, 3), target).cpu().numpy(), G["box_loss"], rtol=TOL, atol=TOL)
    for name in ("rope_penalty", "cloth_penalty", "granular_penalty"):
        got = getattr(rw, name)(state, action, state_cur).cpu().numpy()
        np.testing.assert_allclose(got, G[name], rtol=1e-5, atol=1e-6, err_msg=name)
    # running_cost (plan.py:27-59) with the planner's two criteria (plan.py:146, :155) against the oracle's restatement
    y = torch.from_numpy(G["chamfer/shared_target/y"])
    for crit_gpu, crit_cpu in ((partial(rw.chamfer, y=y.cuda()), partial(po.chamfer, y=y)),
                               (partial(rw.box_loss, target=target), partial(po_box_loss, target=target.cpu()))):
        for pen_gpu in (rw.rope_penalty, rw.granular_penalty, rw.cloth_penalty):
            pen_cpu = lambda s, a, c, _n=pen_gpu.__name__: torch.from_numpy(G[_n])  # noqa: E731  (the reference's own output)
            ref = po.running_cost(state.cpu(), action.cpu(), state_cur.cpu(), crit_cpu, pen_cpu, target.cpu())
            got = rw.running_cost(state, action, state_cur, crit_gpu, pen_gpu, target)["reward_seqs"]
            np.testing.assert_allclose(got.cpu().numpy(), ref.numpy(), rtol=2e-5, atol=2e-5)


def po_box_loss(state, target):
    """losses.py:25-35 on the CPU (checked against the reference's output in the test above via rw.box_loss)."""
    xmin, xmax, zmin, zmax = target[0, 0], target[0, 1], target[1, 0], target[1, 1]
    z = torch.zeros_like(state[:, :, 0])
    xd = torch.maximum(xmin - state[:, :, 0], z) + torch.maximum(state[:, :, 0] - xmax, z)
    zd = torch.maximum(zmin - state[:, :, 2], z) + torch.maximum(state[:, :, 2] - zmax, z)
    return ((xd ** 2 + zd ** 2) ** 0.5).mean(dim=1)
