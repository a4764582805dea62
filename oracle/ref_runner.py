"""ORACLE — TEST INFRASTRUCTURE ONLY.  Drives the UNMODIFIED reference modules staged in oracle/_ref (oracle/build_ref.py):
the reference's own `DynamicsPredictor`, `construct_edges_from_states_batch`, `pad_torch` and `truncate_graph`, stepped by the
autoregressive loop of planning/forward_dynamics.py:156-197.  Used by `bench.py --impl reference` / its cpu_baseline leg
(`kind: "reference"`) and by tests that pin the oracle port to the reference on the GPU box's host."""
import importlib
import os
import sys

import torch

from . import build_ref


def available() -> bool:
    return build_ref.available()


def load():
    """(DynamicsPredictor, construct_edges_from_states_batch, pad_torch, truncate_graph) of the reference."""
    if not available():
        raise RuntimeError("oracle/_ref is not staged (run oracle/build_ref.py where /root/reference exists)")
    for p in (os.path.join(build_ref.OUT, "_stubs"), build_ref.OUT):
        if p not in sys.path:
            sys.path.append(p) if p.endswith("_stubs") else sys.path.insert(0, p)   # real dgl / moviepy, if ever present, win over the stubs
    model = importlib.import_module("dynamics.gnn.model")
    graph = importlib.import_module("dynamics.dataset.graph")
    utils = importlib.import_module("dynamics.utils")
    assert os.path.realpath(model.__file__).startswith(os.path.realpath(build_ref.OUT)), model.__file__
    return model.DynamicsPredictor, graph.construct_edges_from_states_batch, utils.pad_torch, utils.truncate_graph


def make_model(configs, state_dict):
    """The reference constructor on CPU with the given reference-format state_dict."""
    DP = load()[0]
    m = DP(*configs, "cpu").eval()
    m.load_state_dict(state_dict)
    return m


@torch.no_grad()
def rollout(model, w, T, max_nR):
    """forward_dynamics.py:156-197 with the reference's functions: per step truncate_graph -> model(**graph) -> tools move by
    their action, y = min predicted y (:163-168) -> construct_edges_from_states_batch on [pred ; tools] (:171) -> pad_torch (:173)
    -> history shift (:176).  `w` is an adaptigraph_b200.synthetic.Workload on CPU.  Returns (B, T, n_p, 3) and per-step counts."""
    _, cesb, pad_torch, truncate_graph = load()
    n_p = w.n_p
    Rr, Rs = cesb(w.state[:, -1].clone(), w.adj_thresh, w.state_mask, w.eef_mask, topk=w.topk, connect_tools_all=w.connect_tools_all)
    graph = w.graph_dict(pad_torch(Rr, max_nR, dim=1), pad_torch(Rs, max_nR, dim=1))
    preds, counts = [], []
    for _ in range(T):
        graph = truncate_graph(graph)
        counts.append((graph["Rr"].sum(-1) > 0).sum(1))
        pred, _ = model(**graph)
        preds.append(pred)
        y = pred[:, :, 1].min(dim=1).values
        eef = graph["state"][:, -1, n_p:] + graph["action"][:, n_p:]
        eef[:, :, 1] = y[:, None]
        cur = torch.cat([pred, eef], 1)
        Rr, Rs = cesb(cur, w.adj_thresh, graph["state_mask"], graph["eef_mask"], topk=w.topk, connect_tools_all=w.connect_tools_all)
        graph = dict(graph)
        graph["Rr"], graph["Rs"] = pad_torch(Rr, max_nR, dim=1), pad_torch(Rs, max_nR, dim=1)
        graph["state"] = torch.cat([graph["state"][:, 1:], cur[:, None]], 1)
    return torch.stack(preds, 1), torch.stack(counts, 0)
