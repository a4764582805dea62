"""Generate tests/golden/eval_rollout.npz by running the UNMODIFIED reference evaluation rollout (build container only):

    python tests/golden/make_golden_eval.py

Runs the reference's own construct_graph (src/dynamics/rollout/graph.py:233-372) and rollout_from_start_graph
(src/dynamics/rollout/rollout.py:20-148) with its own DynamicsPredictor on the CPU over synthetic episodes (a drifting rope
pushed by one tool point).  Modules absent from the image and never executed here (dgl, moviepy, matplotlib, h5py-backed
sim.data_gen.data) are stubbed; dgl's sampler is the oracle restatement (oracle/sampling_oracle.py).  The model is wrapped only
to RECORD its outputs per step.
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
    from oracle import sampling_oracle as so
    sys.path.insert(0, REF)

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    def sampler(pos, npoints, start_idx=-1):
        return torch.from_numpy(so.farthest_point_sampler(pos.numpy(), npoints, [start_idx] * pos.shape[0]))
    geo = mod("dgl.geometry", farthest_point_sampler=sampler)
    mod("dgl", geometry=geo)
    mod("moviepy", editor=mod("moviepy.editor"))
    plt = mod("matplotlib.pyplot")
    mod("matplotlib", use=lambda *a, **k: None, pyplot=plt)
    mod("sim.data_gen.data", load_data=lambda *a, **k: (_ for _ in ()).throw(RuntimeError("stub")))
    from dynamics.gnn.model import DynamicsPredictor
    from dynamics.rollout.graph import construct_graph, get_next_pair_or_break_episode, get_next_pair_or_break_episode_pushes
    from dynamics.rollout.rollout import rollout_from_start_graph
    return DynamicsPredictor, construct_graph, rollout_from_start_graph, get_next_pair_or_break_episode, get_next_pair_or_break_episode_pushes


def make_episode(seed, n_frames, n_raw, n_eef):
    rng = np.random.default_rng(seed)
    x = np.linspace(0, 3.0, n_raw)
    base = np.stack([x, np.zeros(n_raw), 0.3 * np.sin(2.0 * x + seed)], -1)
    drift = np.stack([0.02 * np.arange(n_frames), np.zeros(n_frames), 0.01 * np.arange(n_frames)], -1)
    obj_pos = (base[None] + drift[:, None] * (0.5 + 0.5 * np.sin(x)[None, :, None]) + rng.normal(0, 0.003, (n_frames, n_raw, 3))).astype(np.float32)
    tool0 = np.array([1.5, 0.05, 0.2]) + 0.08 * np.arange(n_eef)[:, None] * np.array([1.0, 0, 0])
    eef_pos = (tool0[None] + np.stack([0.03 * np.arange(n_frames), np.zeros(n_frames), -0.01 * np.arange(n_frames)], -1)[:, None]).astype(np.float32)
    return eef_pos, obj_pos


def main():
    from adaptigraph_b200 import synthetic as syn
    DP, construct_graph, rollout_from_start_graph, next_episode, next_pushes = import_reference()
    torch.set_num_threads(4)
    model_config, material_config, dataset_config = syn.configs("rope", 2)
    dataset_config = dict(dataset_config)
    dataset_config["datasets"] = [dict(max_nobj=40, max_nR=400, fps_radius_range=[0.12, 0.16], adj_radius_range=[0.3, 0.4], topk=6,
                                       connect_tool_all=False)]
    n_his = dataset_config["n_his"]
    torch.manual_seed(0)
    model = DP(model_config, material_config, dataset_config, "cpu").eval()
    preds = []

    class Recorder(torch.nn.Module):
        def forward(self, **graph):
            out = model(**graph)
            preds.append(out[0][0].detach().numpy().copy())
            return out
    rec = Recorder()

    out = {"n_his": np.int64(n_his)}
    cases = [("stride2", 11, 44, 120, 1, 2, next_pushes, False), ("stride3_gap", 12, 60, 150, 2, 3, next_episode, True)]
    for name, seed, n_frames, n_raw, n_eef, stride, next_fn, drop in cases:
        eef_pos, obj_pos = make_episode(seed, n_frames, n_raw, n_eef)
        rows = [[f + stride * k for k in range(-(n_his - 1), 2)] for f in range(stride * (n_his - 1), n_frames - stride)]
        pairs = np.array(rows, dtype=int)
        if drop:   # holes in the pair list: get_next_pair_or_break_episode walks forward to the next frame that has a pair
            pairs = pairs[(pairs[:, n_his - 1] % 7) != 3]
        pair = pairs[0]
        physics_param = {"rope": np.array([0.35], dtype=np.float32)}
        np.random.seed(seed)
        graph, fps_idx_list = construct_graph(dataset_config, material_config, eef_pos, obj_pos, n_his, pair, physics_param)
        preds.clear()
        errors = rollout_from_start_graph(graph, fps_idx_list, dataset_config, material_config, rec, "cpu", eef_pos, obj_pos,
                                          pair[n_his - 1], pair[n_his], next_fn, pairs, None, False, None, None)
        from tests_agx_helpers import lists_from_onehots  # noqa
        r, s = lists_from_onehots(graph["Rr"][None], graph["Rs"][None])
        for k in ("state", "action", "attrs", "p_rigid", "p_instance", "state_mask", "eef_mask", "obj_mask", "material_index", "eef_kp",
                  "rope_physics_param"):
            out[f"{name}/graph/{k}"] = graph[k].numpy()
        out[f"{name}/graph/recv"], out[f"{name}/graph/send"] = r[0].numpy(), s[0].numpy()
        out[f"{name}/fps_idx_list"] = np.asarray(fps_idx_list, np.int64)
        out[f"{name}/eef_pos"], out[f"{name}/obj_pos"], out[f"{name}/pairs"] = eef_pos, obj_pos, pairs
        out[f"{name}/start_end"] = np.array([pair[n_his - 1], pair[n_his]], np.int64)
        out[f"{name}/pushes"] = np.int64(next_fn is next_pushes)
        out[f"{name}/errors"] = np.asarray(errors, np.float64)
        out[f"{name}/preds"] = np.stack(preds, 0)
        print(name, "steps", len(errors), "particles", len(fps_idx_list), "errors", np.round(errors[:4], 5), "...", np.round(errors[-1], 5))
    out["dataset"] = np.array([40, 400, 0.35, 6, 0], np.float64)   # max_nobj, max_nR, adj_thresh (mean of the range), topk, connect_tool_all
    np.savez_compressed(os.path.join(HERE, "eval_rollout.npz"), **out)


if __name__ == "__main__":
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import agx_helpers
    sys.modules["tests_agx_helpers"] = agx_helpers
    main()
