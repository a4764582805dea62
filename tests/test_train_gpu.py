"""Training glue: fused Adam against torch.optim.Adam (the reference's optimiser, train.py:63), the Trainer's
unroll against the reference's loss (tests/golden/train_unroll_rope40.npz), CUDA-graph replay against eager stepping, and the
optimizer state round trip through torch.optim.Adam's state_dict."""
import copy

import numpy as np
import pytest
import torch

import agx_helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def agx():
    import adaptigraph_b200 as pkg
    import adaptigraph_b200.ops  # noqa: F401  (loads the .so; raises if missing)
    assert torch.cuda.is_available()
    return pkg


def _golden_batch():
    g = H.load_npz("train_unroll_rope40.npz")
    t = lambda k: torch.from_numpy(g[k]).cuda()  # noqa: E731
    N = g["attrs"].shape[1]
    Rr, Rs = H.onehots_from_lists(g["recv"], g["send"], N)
    data = {"state": t("state"), "attrs": t("attrs"), "action": t("action"), "p_instance": t("p_instance"), "Rr": Rr.cuda(), "Rs": Rs.cuda(),
            "rope_physics_param": t("physics_param"), "state_future": t("state_future"), "eef_future": t("eef_future"),
            "action_future": t("action_future")}
    return g, data


def _model(agx, pstep):
    from adaptigraph_b200 import synthetic as syn
    m = agx.DynamicsPredictor(*syn.configs("rope", pstep), "cuda")
    m.load_state_dict(H.golden_weights())
    return m.cuda().train()


def test_fused_adam_matches_torch_adam():
    from adaptigraph_b200 import ops
    torch.manual_seed(0)
    n = 252903                                   # the model's parameter count: exercises the unaligned tail
    p0 = torch.randn(n)
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-3)
    p = p0.cuda()
    m, v, step = torch.zeros_like(p), torch.zeros_like(p), torch.zeros(1, dtype=torch.int32, device="cuda")
    for it in range(4):
        g = torch.randn(n) * (10.0 ** (it - 2))
        ref.grad = g.clone()
        opt.step()
        ops.adam_step(p, (g * 2).cuda(), m, v, step, 1e-3, 0.9, 0.999, 1e-8, grad_scale=0.5)
        assert int(step.item()) == it + 1
        assert (p.cpu() - ref.detach()).abs().max().item() <= 2e-7 * max(1.0, float(ref.detach().abs().max()))
    st = opt.state[ref]
    assert (m.cpu() - st["exp_avg"]).abs().max() <= 1e-6 * st["exp_avg"].abs().max()
    assert (v.cpu() - st["exp_avg_sq"]).abs().max() <= 1e-6 * st["exp_avg_sq"].abs().max()
    with pytest.raises(ValueError):
        ops.adam_step(p, p[:10], m, v, step, 1e-3, 0.9, 0.999, 1e-8)
    with pytest.raises(RuntimeError):
        ops.adam_step(p.cpu(), p.cpu(), m.cpu(), v.cpu(), step.cpu(), 1e-3, 0.9, 0.999, 1e-8)


def test_trainer_matches_reference_loss_and_torch_adam(agx):
    """Same engine gradients, two optimisers: the fused flat-bucket Adam and torch.optim.Adam stepping the reference way."""
    from adaptigraph_b200.train import Trainer, unroll_loss
    g, data = _golden_batch()
    a, b = _model(agx, int(g["pstep"])), _model(agx, int(g["pstep"]))
    tr = Trainer(a, lr=1e-3, n_future=3)
    opt = torch.optim.Adam(b.parameters(), lr=1e-3)
    for it in range(3):
        d = dict(data)
        d["state"] = data["state"] + 0.01 * it
        loss_a = tr.step(d)
        opt.zero_grad()
        loss_b = unroll_loss(b, d, 3)
        loss_b.backward()
        opt.step()
        if it == 0:
            assert abs(loss_a.item() - float(g["loss"])) <= 1e-6          # the reference's own loss on this batch
        assert abs(loss_a.item() - loss_b.item()) <= 1e-6 * max(1.0, abs(loss_b.item()))
        for (k, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
            assert (pa - pb).abs().max().item() <= 5e-6, (it, k)
    assert int(tr.step_count.item()) == 3
    # every parameter is a view of the flat bucket, in model.parameters() order
    assert sum(p.numel() for p in a.parameters()) == tr.bucket.flat.numel() == 252903
    assert next(a.parameters()).data_ptr() == tr.bucket.flat.data_ptr()


def test_cuda_graph_replay_equals_eager_steps(agx):
    from adaptigraph_b200.train import Trainer
    g, data = _golden_batch()
    a, b = _model(agx, int(g["pstep"])), _model(agx, int(g["pstep"]))
    eager, graphed = Trainer(a, n_future=3), Trainer(b, n_future=3, cuda_graph=True)
    edges = agx.edges_from_onehots(data["Rr"], data["Rs"])
    for it in range(4):
        d = dict(data)
        d["state"] = data["state"] + 0.02 * it
        d["state_future"] = data["state_future"] - 0.01 * it
        la, lb = eager.step(d, edges), graphed.step(d, edges)
        assert la.item() == lb.item(), it
    for (k, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
        assert torch.equal(pa, pb), k
    assert int(graphed.step_count.item()) == 4
    bad = dict(data)
    bad["state"] = data["state"][:1]
    with pytest.raises(RuntimeError):
        graphed.step(bad, edges)


def test_cuda_graph_replay_follows_changing_relations(agx):
    """Every batch of a real training run has its own relations: the sender lists the backward reads must be rebuilt inside the
    captured step (they were once taken from a cache filled by the warm-up, i.e. frozen at the first batch's)."""
    from adaptigraph_b200 import synthetic as syn
    from adaptigraph_b200.train import Trainer
    g, data = _golden_batch()
    a, b = _model(agx, int(g["pstep"])), _model(agx, int(g["pstep"]))
    eager, graphed = Trainer(a, n_future=3), Trainer(b, n_future=3, cuda_graph=True)
    B, H_, N, _ = data["state"].shape
    n_rel = data["Rr"].shape[1]
    thr, topk, cta, _ = syn.MATERIALS["rope"]
    mask = torch.ones(B, N, dtype=torch.bool, device="cuda")
    tool = torch.zeros(B, N, dtype=torch.bool, device="cuda")
    tool[:, g["p_instance"].shape[1]:] = True
    seen = set()
    for it in range(4):
        d = dict(data)
        d["state"] = data["state"] * (1.0 + 0.15 * it)                      # stretches the rope: other neighbours inside the radius
        el = agx.build_edges(d["state"][:, -1], thr, mask, tool, topk, cta, max_nR=n_rel).check()
        seen.add(int(el.row_ptr[-1]))
        la, lb = eager.step(d, el), graphed.step(d, el)
        assert la.item() == lb.item(), it
        for (k, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
            assert torch.equal(pa, pb), (it, k)
    assert len(seen) > 1, "the test needs relation sets that differ between steps"


def test_weights_seen_by_inference_follow_the_optimizer(agx):
    """The fused Adam writes the flat bucket through raw pointers: a forward / rollout right after Trainer.step must run on the
    updated weights (the packed blob is dropped), and zero_grad(set_to_none=True) between steps must not detach the bucket."""
    from adaptigraph_b200.train import Trainer
    g, data = _golden_batch()
    a = _model(agx, int(g["pstep"]))
    for graph_mode in (False, True):
        tr = Trainer(a, n_future=3, cuda_graph=graph_mode)
        edges = agx.edges_from_onehots(data["Rr"], data["Rs"])
        for _ in range(2):
            tr.step(data, edges)
            a.zero_grad()                                                    # set_to_none=True by default
        fresh = _model(agx, int(g["pstep"]))
        fresh.load_state_dict({k: v.detach().clone() for k, v in a.state_dict().items()})
        a.eval(); fresh.eval()
        with torch.no_grad():
            kw = {k: v for k, v in data.items() if k not in ("Rr", "Rs")}
            pos_a, _ = a(**kw, edges=edges)
            pos_f, _ = fresh(**kw, edges=edges)
        assert torch.equal(pos_a, pos_f)
        a.train()
    assert not torch.equal(a.state_dict()["particle_encoder.model.0.weight"].cpu(), H.golden_weights()["particle_encoder.model.0.weight"])


def test_optimizer_state_round_trips_through_torch_adam(agx):
    from adaptigraph_b200.train import Trainer, unroll_loss
    g, data = _golden_batch()
    a = _model(agx, int(g["pstep"]))
    tr = Trainer(a, n_future=3)
    assert tr.optimizer_state_dict()["state"] == {}
    tr.step(data)
    tr.step(data)
    sd = tr.optimizer_state_dict()
    b = _model(agx, int(g["pstep"]))
    b.load_state_dict(copy.deepcopy(a.state_dict()))
    opt = torch.optim.Adam(b.parameters(), lr=1e-3)
    opt.load_state_dict(sd)                                   # the reference's latest_optim.pth format (train.py:121)
    tr.step(data)
    opt.zero_grad()
    unroll_loss(b, data, 3).backward()
    opt.step()
    for (k, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
        assert (pa - pb).abs().max().item() <= 5e-6, k
    c = _model(agx, int(g["pstep"]))
    tr2 = Trainer(c, n_future=3)
    tr2.load_optimizer_state_dict(opt.state_dict())
    assert int(tr2.step_count.item()) == 3
    ref_m = opt.state_dict()["state"][0]["exp_avg"]
    assert torch.equal(tr2.bucket.views(tr2.exp_avg)[0], ref_m.cuda())


class _ReluWithPattern(torch.autograd.Function):
    """relu(x) whose backward uses a given 0/1 activity pattern instead of x > 0."""

    @staticmethod
    def forward(ctx, x, pattern):
        ctx.save_for_backward(pattern)
        return x.clamp_min(0)

    @staticmethod
    def backward(ctx, g):
        (pattern,) = ctx.saved_tensors
        return g * pattern, None


def _engine_patterns(agx, m, wd, el, pstep):
    """ReLU activity patterns of the engine's training forward (agx_forward_train run once more on the same inputs: it is
    deterministic), keyed like the sites of oracle.forward_sparse."""
    from adaptigraph_b200 import ops
    B, H_, N, _ = wd.state.shape
    n_p = wd.p_instance.shape[1]
    f32 = lambda t: t.to(torch.float32).contiguous()  # noqa: E731
    p_inst = f32(wd.p_instance.reshape(B, n_p, -1)[:, :, 0])
    F = m.nf_effect
    _, _, saved = ops.forward_train(m.packed_weights(), f32(wd.state), f32(wd.attrs), f32(wd.action), p_inst, f32(wd.physics_param),
                                    el.row_ptr, el.send, el.recv, F, pstep)
    v = ops.train_saved_views(saved, F, H_, wd.attrs.shape[2], wd.physics_param.shape[1], wd.action.shape[2], pstep, B, N, int(el.send.numel()))
    E = int(el.row_ptr[-1])
    recv = el.recv[:E].long()
    snd = el.send[:E].long() + (recv // N) * N
    pat = {k: (v[k][:, :F] > 0) for k in ("h1", "h2", "penc")}
    pat.update({k: (v[k][:E, :F] > 0) for k in ("g1", "g2", "renc")})
    for k in range(pstep):
        pat[f"e{k}"] = ((v["C"][:E, :F] + v[f"Qr{k}"][recv, :F]) + v[f"Qs{k}"][snd, :F]) > 0      # the backward's own association
        pat[f"P{k + 1}"] = v[f"P{k + 1}"][:, :F] > 0
    for k in ("u1", "u2"):
        pat[k] = v[k].view(B, N, -1)[:, :n_p, :F] > 0
    return {k: t.cpu() for k, t in pat.items()}, E


@pytest.mark.parametrize("material,n_p,B,pstep", [("granular", 150, 3, 2), ("cloth", 200, 2, 3), ("rope", 120, 3, 4)])
def test_gradients_match_oracle_autograd_on_seeded_graphs(agx, material, n_p, B, pstep):
    """Beyond the golden unroll (rope, 40 particles): loss and every gradient of one forward + backward (tensor-core chains)
    against torch autograd over the oracle, on graphs with hundreds of particles, tool senders and padded particles.

    A gradient is a discontinuous function of the forward values: wherever a ReLU input is within rounding of zero, two fp32
    evaluations of the same model (torch on the CPU, torch on a GPU, this engine) may disagree on whether the unit is active, and
    each such unit moves the gradients by up to 1e-4 of their maximum.  The comparison is therefore stated the way the rollout
    parity is stated for relation flips: (a) the engine's activity pattern differs from the oracle's only at
    ReLU inputs the oracle itself puts within 1e-5 of zero (a handful among millions); (b) on the SAME activity pattern -- the
    oracle's autograd with the engine's pattern substituted in the ReLU backward -- every gradient agrees to 1e-5 of its maximum;
    (c) against the oracle's unconditioned gradients the difference stays below 3e-4."""
    from adaptigraph_b200 import synthetic as syn
    from oracle import dynamics_oracle as orc
    w = syn.make_workload(material, n_p, B, seed=77, n_pad=6)
    torch.manual_seed(3)
    m = agx.DynamicsPredictor(*syn.configs(material, pstep), "cuda").cuda().train()
    wd = w.to("cuda")
    el = agx.build_edges(wd.state[:, -1], w.adj_thresh, wd.state_mask, wd.eef_mask, w.topk, w.connect_tools_all).check()
    st = wd.state.clone().requires_grad_(True)
    d = wd.graph_dict()
    d["state"] = st
    pos, motion = m(**d, edges=el)
    tgt = torch.randn(pos.shape, generator=torch.Generator().manual_seed(5)).cuda() * 0.05
    loss = torch.nn.functional.mse_loss(pos, wd.state[:, -1, :pos.shape[1]] + tgt) + 0.1 * motion.square().mean()
    loss.backward()
    patterns, E = _engine_patterns(agx, m, wd, el, pstep)

    def oracle_grads(relu):
        p = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.state_dict().items()}
        st_c = w.state.clone().requires_grad_(True)
        pos_c, motion_c = orc.forward_sparse(p, pstep, st_c, w.attrs, el.row_ptr.cpu(), el.send.cpu()[:E], w.p_instance, w.action,
                                             w.physics_param, relu=relu)
        loss_c = torch.nn.functional.mse_loss(pos_c, w.state[:, -1, :pos_c.shape[1]] + tgt.cpu()) + 0.1 * motion_c.square().mean()
        loss_c.backward()
        return loss_c, st_c.grad, {k: v.grad for k, v in p.items()}

    # (a) where the patterns differ, and (b) the oracle's gradients on the engine's pattern
    differing = []

    def with_engine_pattern(site, x):
        pat = patterns[site].reshape(x.shape)
        diff = pat != (x.detach() > 0)
        if diff.any():
            differing.append((site, int(diff.sum()), float(x.detach()[diff].abs().max())))
        return _ReluWithPattern.apply(x, pat.to(x.dtype))

    loss_c, st_ref, p_ref = oracle_grads(with_engine_pattern)
    assert abs(loss.item() - loss_c.item()) <= 1e-6 * max(1.0, abs(loss_c.item()))
    assert sum(n for _, n, _ in differing) <= 64, differing
    assert all(mx <= 1e-5 for _, _, mx in differing), differing
    assert (st.grad.cpu() - st_ref).abs().max() <= 1e-5 * max(1e-3, float(st_ref.abs().max()))
    for k, v in m.named_parameters():
        assert (v.grad.cpu() - p_ref[k]).abs().max() <= 1e-5 * max(1e-3, float(p_ref[k].abs().max())), k
    # (c) the unconditioned oracle
    _, st_raw, p_raw = oracle_grads(None)
    assert (st.grad.cpu() - st_raw).abs().max() <= 3e-4 * max(1e-3, float(st_raw.abs().max()))
    for k, v in m.named_parameters():
        assert (v.grad.cpu() - p_raw[k]).abs().max() <= 3e-4 * max(1e-3, float(p_raw[k].abs().max())), k
