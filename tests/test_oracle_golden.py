"""Pins oracle/dynamics_oracle.py against outputs of the reference itself
(tests/golden/*.npz, produced by tests/golden/make_golden.py from /root/reference)."""
import numpy as np
import pytest
import torch

from oracle import dynamics_oracle as orc
import agx_helpers as H

CASES = H.load_npz("graph_cases.npz")
TOL = 2e-6


@pytest.mark.parametrize("name", H.graph_case_names(CASES))
def test_graph_builder_batch_matches_reference(name):
    c = {k.split("/", 1)[1]: v for k, v in CASES.items() if k.startswith(name + "/")}
    pos, mask, tool = torch.from_numpy(c["pos"]), torch.from_numpy(c["mask"]), torch.from_numpy(c["tool_mask"])
    thr = torch.from_numpy(c["adj_thresh"]) if bool(c["thr_is_tensor"]) else float(c["adj_thresh"])
    Rr, Rs = orc.edges_dense_batch(pos, thr, mask, tool, int(c["topk"]), bool(c["cta"]))
    r, s = H.lists_from_onehots(Rr, Rs)
    assert np.array_equal(r.numpy(), c["batch_recv"]) and np.array_equal(s.numpy(), c["batch_send"])
    # CSR form agrees with the dense form
    adj = orc.adjacency_batch(pos, thr, mask, tool, int(c["topk"]), bool(c["cta"]))
    row_ptr, send = orc.edge_lists_from_adjacency(adj)
    rp2, s2 = H.csr_from_lists(c["batch_recv"], c["batch_send"], pos.shape[1])
    assert torch.equal(row_ptr, rp2) and torch.equal(send, s2)


@pytest.mark.parametrize("name", [n for n in H.graph_case_names(CASES) if f"{n}/single_recv_0" in CASES])
def test_graph_builder_single_matches_reference(name):
    c = {k.split("/", 1)[1]: v for k, v in CASES.items() if k.startswith(name + "/")}
    pos, mask, tool = torch.from_numpy(c["pos"]), torch.from_numpy(c["mask"]), torch.from_numpy(c["tool_mask"])
    for b in range(pos.shape[0]):
        Rr, Rs = orc.edges_dense_single(pos[b], float(c["adj_thresh"]), mask[b], tool[b], int(c["topk"]), bool(c["cta"]))
        r, s = H.lists_from_onehots(Rr[None], Rs[None])
        assert np.array_equal(r[0].numpy(), c[f"single_recv_{b}"]) and np.array_equal(s[0].numpy(), c[f"single_send_{b}"])


FWD = ["forward_rope100_k1.npz", "forward_cloth64_pad_k3.npz", "forward_granular120_k3.npz", "forward_rope300_k4.npz"]


def _fwd_inputs(g):
    t = lambda k: torch.from_numpy(g[k])  # noqa: E731
    return t("state"), t("attrs"), t("p_instance"), t("action"), t("physics_param")


@pytest.mark.parametrize("fname", FWD)
def test_forward_dense_and_sparse_match_reference(fname):
    g = H.load_npz(fname)
    p = H.golden_weights()
    state, attrs, p_inst, action, phys = _fwd_inputs(g)
    N = attrs.shape[1]
    Rr, Rs = H.onehots_from_lists(g["recv"], g["send"], N)
    pos, motion = orc.forward_dense(p, int(g["pstep"]), state, attrs, Rr, Rs, p_inst, action, phys)
    assert np.abs(pos.numpy() - g["pred_pos"]).max() <= TOL
    assert np.abs(motion.numpy() - g["pred_motion"]).max() <= TOL
    row_ptr, send = H.csr_from_lists(g["recv"], g["send"], N)
    pos2, motion2 = orc.forward_sparse(p, int(g["pstep"]), state, attrs, row_ptr, send, p_inst, action, phys)
    assert np.abs(pos2.numpy() - g["pred_pos"]).max() <= TOL
    assert np.abs(motion2.numpy() - g["pred_motion"]).max() <= TOL


def test_padded_relation_rows_do_not_change_outputs():
    g = H.load_npz("forward_cloth64_pad_k3.npz")
    p = H.golden_weights()
    state, attrs, p_inst, action, phys = _fwd_inputs(g)
    Rr, Rs = H.onehots_from_lists(g["recv"], g["send"], attrs.shape[1])
    assert Rr.shape[1] == 600
    Rr_t, Rs_t = orc.truncate_rows(Rr, Rs)
    assert Rr_t.shape[1] < 600
    a, _ = orc.forward_dense(p, 3, state, attrs, Rr, Rs, p_inst, action, phys)
    b, _ = orc.forward_dense(p, 3, state, attrs, Rr_t, Rs_t, p_inst, action, phys)
    assert (a - b).abs().max() <= TOL
    with pytest.raises(RuntimeError):
        orc.pad_rows(Rr, 10)


@pytest.mark.parametrize("fname", ["rollout_rope60_T10.npz", "rollout_cloth49_T10.npz", "rollout_granular80_T5.npz"])
def test_rollout_matches_reference(fname):
    from adaptigraph_b200 import synthetic as syn
    g = H.load_npz(fname)
    p = H.golden_weights()
    state, attrs, p_inst, action, phys = _fwd_inputs(g)
    mat = str(g["material"])
    thr, topk, cta, _ = syn.MATERIALS[mat]
    T = g["preds"].shape[1]
    preds, edges = orc.rollout_dense(p, int(g["pstep"]), state, attrs, p_inst, action, phys,
                                     torch.from_numpy(g["mask"]), torch.from_numpy(g["tool_mask"]), thr, topk, cta, T)
    # edge sets identical at every step, positions within fp32 noise
    for t, (Rr, Rs) in enumerate(edges):
        r, s = H.lists_from_onehots(Rr, Rs)
        n = r.shape[1]
        assert np.array_equal(r.numpy(), g["recv"][:, t, :n]) and np.array_equal(s.numpy(), g["send"][:, t, :n])
        assert (g["recv"][:, t, n:] == -1).all()
    assert np.abs(preds.numpy() - g["preds"]).max() <= 1e-5


def test_training_unroll_loss_and_grads_match_reference():
    g = H.load_npz("train_unroll_rope40.npz")
    p = {k: v.clone().requires_grad_(True) for k, v in H.golden_weights().items()}
    state, attrs, p_inst, action, phys = _fwd_inputs(g)
    state = state.clone().requires_grad_(True)
    Rr, Rs = H.onehots_from_lists(g["recv"], g["send"], attrs.shape[1])
    t = lambda k: torch.from_numpy(g[k])  # noqa: E731
    loss = orc.unroll_loss_dense(p, int(g["pstep"]), state, attrs, Rr, Rs, p_inst, action, phys,
                                 t("state_future"), t("eef_future"), t("action_future"))
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= 1e-6
    assert np.abs(state.grad.numpy() - g["grad_state"]).max() <= 1e-6
    for k, v in p.items():
        ref = g["grad/" + k]
        assert np.abs(v.grad.numpy() - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max()), k


def test_init_matches_reference_constructor_order():
    """torch.manual_seed(0) + nn.Linear construction order of model.py:103-122 reproduces the golden weights."""
    import torch.nn as nn
    torch.manual_seed(0)
    dims = [(6, 150), (150, 150), (150, 150), (17, 150), (150, 150), (150, 150), (300, 150), (450, 150),
            (150, 150), (150, 150), (150, 3)]
    layers = [nn.Linear(i, o) for i, o in dims]
    gw = H.golden_weights()
    names = list(orc.PARAM_SHAPES(6, 17, 150).keys())
    for lyr, nm in zip(layers, names):
        assert torch.equal(lyr.weight.detach(), gw[nm + ".weight"]), nm


@pytest.mark.parametrize("name", H.graph_case_names(CASES))
def test_c_graph_oracle_matches_reference(name):
    from oracle import build_oracle
    c = {k.split("/", 1)[1]: v for k, v in CASES.items() if k.startswith(name + "/")}
    thr = np.float32(c["adj_thresh"])
    B, N = c["mask"].shape
    recv, send, n_edges = build_oracle.edges(c["pos"], c["mask"], c["tool_mask"], thr * thr, int(c["topk"]), bool(c["cta"]), 0)
    off = 0
    for b in range(B):
        n = int(n_edges[b])
        assert n == int((c["batch_recv"][b] >= 0).sum())
        assert np.array_equal(recv[off:off + n], c["batch_recv"][b, :n]) and np.array_equal(send[off:off + n], c["batch_send"][b, :n])
        off += n
    if f"single_recv_0" in c:
        t2 = np.float32(float(c["adj_thresh"]) * float(c["adj_thresh"]))
        for b in range(B):
            r1, s1, _ = build_oracle.edges(c["pos"][b:b + 1], c["mask"][b:b + 1], c["tool_mask"][b:b + 1], t2, int(c["topk"]), bool(c["cta"]), 1)
            assert np.array_equal(r1, c[f"single_recv_{b}"]) and np.array_equal(s1, c[f"single_send_{b}"])


PLAN = H.load_npz("planning_dynamics.npz")


def _plan_cfg(name):
    from adaptigraph_b200 import synthetic as syn
    c = {k.split("/", 1)[1]: v for k, v in PLAN.items() if k.startswith(name + "/")}
    thr, topk, cta, _ = syn.MATERIALS[str(c["material"])]
    cfg = dict(pusher=c["pusher"].tolist(), ratio=10.0, push_length=0.1, gripper=bool(c["gripper"]), thr=thr, topk=topk, cta=cta, n_his=4, phys=0.4)
    return c, cfg


@pytest.mark.parametrize("name", ["rope1pt", "granular5pt", "cloth_gripper"])
def test_mpc_drivers_match_reference(name):
    """forward_dynamics.py `dynamics` and `dynamics_masked` (reference outputs) vs the oracle restatement."""
    from oracle import planning_oracle as po
    c, cfg = _plan_cfg(name)
    p = H.golden_weights()
    seq, dec = po.dynamics(p, 3, torch.from_numpy(c["state"]), torch.from_numpy(c["action"]), cfg)
    assert np.abs(dec.numpy() - c["action_seqs"]).max() <= 1e-6
    assert np.abs(seq.numpy() - c["state_seqs"]).max() <= 2e-5
    seq_m, dec_m = po.dynamics_masked(p, 3, torch.from_numpy(c["m_state"]), torch.from_numpy(c["m_mask"]), torch.from_numpy(c["action"][:, 0]), cfg)
    assert np.abs(dec_m.numpy() - c["m_action_seqs"]).max() <= 1e-6
    assert np.abs(seq_m.numpy() - c["m_state_seqs"]).max() <= 2e-5
