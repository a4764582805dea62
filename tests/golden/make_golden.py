"""Generate tests/golden/*.npz by running the UNMODIFIED reference from /root/reference.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

`dgl` and `moviepy` are absent from the image and are imported at module top by
dynamics/dataset/graph.py:5 and dynamics/utils.py:6,8 although never used on the
hot path, so two empty stub modules are registered before the import
(SURVEY.md §8c).  Everything else is the reference's own code: its
DynamicsPredictor constructor (weights under torch.manual_seed(0)), its
construct_edges_from_states[_batch], pad_torch and truncate_graph.

Dense Rr/Rs are stored as edge lists (receiver, sender per relation row, -1 for
zero-padded rows) to keep fixtures small.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

REF = "/root/reference/src"


def import_reference():
    sys.path.insert(0, REF)
    dgl = types.ModuleType("dgl")
    geo = types.ModuleType("dgl.geometry")
    geo.farthest_point_sampler = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("stub"))
    dgl.geometry = geo
    mp = types.ModuleType("moviepy")
    mpe = types.ModuleType("moviepy.editor")
    mp.editor = mpe
    sys.modules.update({"dgl": dgl, "dgl.geometry": geo, "moviepy": mp, "moviepy.editor": mpe})
    from dynamics.gnn.model import DynamicsPredictor
    from dynamics.dataset.graph import construct_edges_from_states, construct_edges_from_states_batch
    from dynamics.utils import pad_torch, truncate_graph
    return DynamicsPredictor, construct_edges_from_states, construct_edges_from_states_batch, pad_torch, truncate_graph


def rel_lists(Rr, Rs):
    """(B, n_rel, N) one-hots -> int32 (B, n_rel) receiver / sender ids, -1 on zero rows."""
    def one(R):
        idx = R.argmax(-1).to(torch.int32)
        idx[R.sum(-1) == 0] = -1
        return idx.numpy()
    return one(Rr), one(Rs)


def main():
    from adaptigraph_b200 import synthetic as syn
    DP, ces, cesb, pad_torch, truncate_graph = import_reference()
    torch.set_num_threads(4)

    def ref_model(material, pstep):
        mc, matc, dc = syn.configs(material, pstep)
        torch.manual_seed(0)
        return DP(mc, matc, dc, "cpu").eval()

    # ---- weights: the reference constructor under manual_seed(0), nf=150 (identical for all materials)
    m = ref_model("rope", 1)
    sd = {k: v.detach().numpy() for k, v in m.state_dict().items()}
    np.savez(os.path.join(HERE, "weights_seed0.npz"), **sd)
    print("weights", sum(v.size for v in sd.values()))

    # ---- graph builder cases (A1 batched, A2 single)
    cases = {}

    def add_graph_case(name, w, adj_thresh=None, topk=None, cta=None, pos=None):
        pos = w.state[:, -1] if pos is None else pos
        thr = w.adj_thresh if adj_thresh is None else adj_thresh
        topk = w.topk if topk is None else topk
        cta = w.connect_tools_all if cta is None else cta
        Rr, Rs = cesb(pos.clone(), thr, w.state_mask, w.eef_mask, topk=topk, connect_tools_all=cta)
        r, s = rel_lists(Rr, Rs)
        d = {"pos": pos.numpy(), "mask": w.state_mask.numpy(), "tool_mask": w.eef_mask.numpy(),
             "adj_thresh": np.asarray(thr.numpy() if torch.is_tensor(thr) else thr, dtype=np.float32),
             "thr_is_tensor": np.asarray(torch.is_tensor(thr)),
             "topk": np.asarray(topk), "cta": np.asarray(cta), "batch_recv": r, "batch_send": s}
        # A2 on every graph of the batch (float threshold only)
        if not torch.is_tensor(thr):
            for b in range(pos.shape[0]):
                Rr1, Rs1 = ces(pos[b].clone(), thr, w.state_mask[b], w.eef_mask[b], topk=topk, connect_tools_all=cta)
                r1, s1 = rel_lists(Rr1[None], Rs1[None])
                d[f"single_recv_{b}"], d[f"single_send_{b}"] = r1[0], s1[0]
        for k, v in d.items():
            cases[f"{name}/{k}"] = v
        print(name, "n_rel", Rr.shape[1])

    add_graph_case("rope100", syn.make_workload("rope", 100, 1, 1235))
    add_graph_case("rope37_pad", syn.make_workload("rope", 37, 3, 11, n_pad=5))
    add_graph_case("granular150", syn.make_workload("granular", 150, 2, 12))
    add_graph_case("granular90_pad_cta", syn.make_workload("granular", 90, 2, 13, n_pad=7), cta=True)
    add_graph_case("cloth64", syn.make_workload("cloth", 64, 3, 14))
    add_graph_case("cloth100_pad", syn.make_workload("cloth", 100, 2, 15, n_pad=19))
    # far tools: no tool-receiver adjacency survives in graph 1 -> batch_mask False there (graph.py:135)
    w = syn.make_workload("cloth", 49, 3, 16)
    pos = w.state[:, -1].clone()
    pos[1, w.n_p:] += 50.0
    add_graph_case("cloth49_fartool", w, pos=pos)
    # per-graph tensor threshold (graph.py:106-108 else-branch)
    w = syn.make_workload("rope", 50, 4, 17)
    add_graph_case("rope50_thr_tensor", w, adj_thresh=torch.tensor([0.3, 0.5, 0.7, 1.1]))
    # topk larger than N, topk == 1, no tools
    add_graph_case("rope6_topk_gt_n", syn.make_workload("rope", 6, 2, 18), topk=10)
    add_graph_case("granular40_topk1", syn.make_workload("granular", 40, 1, 19), topk=1)
    add_graph_case("cloth36_notool", syn.make_workload("cloth", 36, 2, 20, n_s=0))
    add_graph_case("cloth36_cta_off", syn.make_workload("cloth", 36, 2, 21), cta=False)
    add_graph_case("rope300", syn.make_workload("rope", 300, 2, 1236))
    np.savez_compressed(os.path.join(HERE, "graph_cases.npz"), **cases)

    # ---- forward cases (dense reference forward; hooks capture per-stage tensors)
    def forward_case(name, w, pstep, pad_to=None, stages=False):
        model = ref_model(w.material, pstep)
        Rr, Rs = cesb(w.state[:, -1].clone(), w.adj_thresh, w.state_mask, w.eef_mask,
                      topk=w.topk, connect_tools_all=w.connect_tools_all)
        if pad_to:
            Rr, Rs = pad_torch(Rr, pad_to, dim=1), pad_torch(Rs, pad_to, dim=1)
        cap = {}
        hooks = []
        if stages:
            hooks.append(model.particle_encoder.register_forward_hook(lambda m, i, o: cap.__setitem__("particle_encode", o.detach().numpy())))
            hooks.append(model.relation_encoder.register_forward_hook(lambda m, i, o: cap.__setitem__("relation_encode", o.detach().numpy())))
            eff = []
            hooks.append(model.particle_propagator.register_forward_hook(lambda m, i, o: eff.append(o.detach().numpy())))
        with torch.no_grad():
            pos, motion = model(**w.graph_dict(Rr, Rs))
        for h in hooks:
            h.remove()
        if stages:
            cap["particle_effect"] = np.stack(eff, 0)
        r, s = rel_lists(Rr, Rs)
        d = {"state": w.state.numpy(), "attrs": w.attrs.numpy(), "action": w.action.numpy(),
             "p_instance": w.p_instance.numpy(), "physics_param": w.physics_param.numpy(),
             "mask": w.state_mask.numpy(), "tool_mask": w.eef_mask.numpy(),
             "recv": r, "send": s, "pstep": np.asarray(pstep), "pred_pos": pos.numpy(), "pred_motion": motion.numpy(),
             "material": np.asarray(w.material), **cap}
        np.savez_compressed(os.path.join(HERE, f"forward_{name}.npz"), **d)
        print("forward", name, "n_rel", Rr.shape[1], "motion rms", float(motion.pow(2).mean().sqrt()))

    forward_case("rope100_k1", syn.baseline_workload(1), 1, stages=True)                 # BASELINE cfg1
    forward_case("cloth64_pad_k3", syn.make_workload("cloth", 64, 3, 31, n_pad=9), 3, pad_to=600, stages=True)
    forward_case("granular120_k3", syn.make_workload("granular", 120, 2, 32), 3)
    forward_case("rope300_k4", syn.make_workload("rope", 300, 2, 1236), 4)               # cfg2 shape, B=2

    # ---- rollout (forward_dynamics.py:156-197 loop, reference functions, re-graph every step)
    def rollout_case(name, w, pstep, T, max_nR):
        model = ref_model(w.material, pstep)
        n_p = w.n_p
        Rr, Rs = cesb(w.state[:, -1].clone(), w.adj_thresh, w.state_mask, w.eef_mask, topk=w.topk, connect_tools_all=w.connect_tools_all)
        graph = w.graph_dict(pad_torch(Rr, max_nR, dim=1), pad_torch(Rs, max_nR, dim=1))
        preds, recvs, sends = [], [], []
        with torch.no_grad():
            for _ in range(T):
                graph = truncate_graph(graph)
                r, s = rel_lists(graph["Rr"], graph["Rs"])
                rp = np.full((w.B, max_nR), -1, np.int32); rp[:, :r.shape[1]] = r
                sp = np.full((w.B, max_nR), -1, np.int32); sp[:, :s.shape[1]] = s
                recvs.append(rp); sends.append(sp)
                pred, _ = model(**graph)
                preds.append(pred.numpy())
                y = pred[:, :, 1].min(dim=1).values
                eef = graph["state"][:, -1, n_p:] + graph["action"][:, n_p:]
                eef[:, :, 1] = y[:, None]
                cur = torch.cat([pred, eef], 1)
                Rr, Rs = cesb(cur, w.adj_thresh, graph["state_mask"], graph["eef_mask"], topk=w.topk, connect_tools_all=w.connect_tools_all)
                graph = dict(graph)
                graph["Rr"], graph["Rs"] = pad_torch(Rr, max_nR, dim=1), pad_torch(Rs, max_nR, dim=1)
                graph["state"] = torch.cat([graph["state"][:, 1:], cur[:, None]], 1)
        d = {"state": w.state.numpy(), "attrs": w.attrs.numpy(), "action": w.action.numpy(),
             "p_instance": w.p_instance.numpy(), "physics_param": w.physics_param.numpy(),
             "mask": w.state_mask.numpy(), "tool_mask": w.eef_mask.numpy(), "pstep": np.asarray(pstep),
             "material": np.asarray(w.material), "preds": np.stack(preds, 1),
             "recv": np.stack(recvs, 1), "send": np.stack(sends, 1)}
        np.savez_compressed(os.path.join(HERE, f"rollout_{name}.npz"), **d)
        print("rollout", name, "step-to-step rms move", float(np.sqrt(((d['preds'][:, -1] - d['preds'][:, 0]) ** 2).mean())))

    rollout_case("rope60_T10", syn.make_workload("rope", 60, 2, 41), 3, 10, 700)
    rollout_case("cloth49_T10", syn.make_workload("cloth", 49, 2, 42), 3, 10, 600)
    rollout_case("granular80_T5", syn.make_workload("granular", 80, 2, 43), 3, 5, 1800)

    # ---- training unroll with autograd (train.py:90-112): loss, parameter grads, grad wrt state
    w = syn.make_workload("rope", 40, 2, 51)
    model = ref_model("rope", 4).train()
    Rr, Rs = cesb(w.state[:, -1].clone(), w.adj_thresh, w.state_mask, w.eef_mask, topk=w.topk, connect_tools_all=False)
    Rr, Rs = pad_torch(Rr, 300, dim=1), pad_torch(Rs, 300, dim=1)
    g = torch.Generator().manual_seed(52)
    n_future = 3
    state_future = w.state[:, -1:, :w.n_p].repeat(1, n_future, 1, 1) + 0.05 * torch.randn(w.B, n_future, w.n_p, 3, generator=g)
    eef_future = w.state[:, -1:].repeat(1, n_future - 1, 1, 1) + w.action[:, None] * torch.arange(1, n_future)[None, :, None, None]
    action_future = w.action[:, None].repeat(1, n_future - 1, 1, 1)
    data = w.graph_dict(Rr, Rs)
    data["state"] = data["state"].clone().requires_grad_(True)
    state_leaf = data["state"]
    loss_sum = 0
    for fi in range(n_future):
        gt = state_future[:, fi].clone()
        pred, _ = model(**data)
        pred_p = pred[:, :gt.shape[1], :3].clone()
        loss_sum = loss_sum + torch.nn.functional.mse_loss(pred_p, gt)
        if fi < n_future - 1:
            nxt = eef_future[:, fi].clone().unsqueeze(1)
            nxt[:, -1, :pred_p.shape[1]] = pred_p
            data["state"] = torch.cat([data["state"][:, 1:], nxt], dim=1)
            data["action"] = action_future[:, fi].clone()
    loss_sum.backward()
    r, s = rel_lists(Rr, Rs)
    d = {"state": w.state.numpy(), "attrs": w.attrs.numpy(), "action": w.action.numpy(),
         "p_instance": w.p_instance.numpy(), "physics_param": w.physics_param.numpy(),
         "recv": r, "send": s, "pstep": np.asarray(4), "state_future": state_future.numpy(),
         "eef_future": eef_future.numpy(), "action_future": action_future.numpy(),
         "loss": np.asarray(loss_sum.item()), "grad_state": state_leaf.grad.numpy()}
    for k, v in model.named_parameters():
        d["grad/" + k] = v.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "train_unroll_rope40.npz"), **d)
    print("train loss", loss_sum.item())


if __name__ == "__main__":
    main()
