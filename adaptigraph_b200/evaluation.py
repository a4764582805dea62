"""Batched evaluation rollout, replacing the batch-1, numpy-round-trip loop of
src/dynamics/rollout/rollout.py:20-148, and the drivers around it (construct_graph, rollout_episode_pushes, rollout_dataset,
rollout: rollout/graph.py:233-372, rollout.py:150-310) reading the reference's on-disk formats through dataset.load_*.

    rollout_from_start_graph(graph, fps_idx_list, dataset_config, material_config, model, device, eef_pos, obj_pos,
                             current_start, current_end, get_next_pair_or_break_func, pairs, ...)   -> error_list
        the reference's signature for ONE episode (visualisation is out of scope: viz=True raises);
    rollout_episodes(model, episodes, dataset_config, get_next_pair_or_break_func)                   -> [error_list, ...]
        any number of episodes advanced TOGETHER on the device.

What the reference does per step — forward, error against the recorded particles, tool points taken from the recording,
action = tool displacement, single-graph relation rebuild (graph.py:38-89) padded to max_nR, history shift — depends on the
prediction only through device tensors, while the frame schedule (which pair comes next, when the episode ends) depends only
on `pairs`.  So the schedule of every episode is walked on the host up front with the caller's get_next_pair_or_break_func,
the recorded tool / ground-truth frames it visits are uploaded once, and the loop itself never leaves the GPU: no .cpu() per
step, no dense one-hots, one host synchronisation at the end (capacity check + error read-back).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Sequence

import numpy as np
import torch

from . import _lib as L
from . import ops
from .graph import EdgeList, edges_from_onehots

_MAX_STEPS = 100   # rollout.py:63


def _schedule(pairs, n_his, n_frames, current_start, current_end, next_fn) -> List[List[int]]:
    """The (start, end) frame pair of every rollout step: rollout.py:65-66, :92-96."""
    idx_list = [[int(current_start), int(current_end)]]
    for _ in range(1, _MAX_STEPS):
        next_pair = next_fn(pairs, n_his, n_frames, idx_list[-1][1])
        if next_pair is None:
            break
        idx_list.append([int(next_pair[n_his - 1]), int(next_pair[n_his])])
    # step i predicts frame idx_list[i-1][1]; the loop runs once more after the last pair was found only if a next pair
    # exists, so the number of forwards equals len(idx_list) unless the step limit cut it
    return idx_list


def rollout_episodes(model, episodes: Sequence[Dict], dataset_config: Dict, get_next_pair_or_break_func: Callable,
                     return_predictions: bool = False):
    """episodes: dicts with the arguments the reference passes per episode — 'graph' (construct_graph's dict, CPU or CUDA
    tensors, unbatched), 'fps_idx_list', 'eef_pos' (T, n_eef, 3), 'obj_pos' (T, n_all, 3), 'current_start', 'current_end',
    'pairs'.  All episodes share the dataset's max_nobj / eef count.  Returns one error list per episode (python floats); with
    return_predictions also the predicted particles (E, T, max_nobj, 3) on the device (rows past an episode's end are zero)."""
    dataset = dataset_config["datasets"][0]
    max_nobj, max_nR = dataset["max_nobj"], dataset["max_nR"]
    adj_thresh = (dataset["adj_radius_range"][0] + dataset["adj_radius_range"][1]) / 2          # rollout.py:34
    topk, connect_tool_all = dataset["topk"], dataset["connect_tool_all"]
    n_his = dataset_config["n_his"]
    dev = next(model.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("rollout_episodes runs on a CUDA model only (no CPU path)")
    E = len(episodes)

    # ---- host: frame schedules, then ONE upload of everything the loop reads
    scheds = []
    for ep in episodes:
        n_frames = ep["obj_pos"].shape[0]
        assert ep["eef_pos"].shape[0] == n_frames                                               # rollout.py:40
        scheds.append(_schedule(ep["pairs"], n_his, n_frames, ep["current_start"], ep["current_end"], get_next_pair_or_break_func))
    T = max(len(s) for s in scheds)
    n_eef = episodes[0]["eef_pos"].shape[1]
    gt = np.zeros((E, T, max_nobj, 3), np.float32)          # recorded particles at the step's end frame, fps-selected, zero padded
    eef_s = np.zeros((E, T, n_eef, 3), np.float32)          # tool points at the NEXT step's start / end frames
    eef_e = np.zeros((E, T, n_eef, 3), np.float32)
    live = np.zeros((E, T), bool)
    for e, (ep, sched) in enumerate(zip(episodes, scheds)):
        fps_idx = np.asarray(ep["fps_idx_list"])
        for i, (_, end) in enumerate(sched):
            g = ep["obj_pos"][end][fps_idx]
            gt[e, i, :g.shape[0]] = g                                                           # pad(): rollout.py:76-78
            live[e, i] = True
            if i + 1 < len(sched):
                eef_s[e, i] = ep["eef_pos"][sched[i + 1][0]]
                eef_e[e, i] = ep["eef_pos"][sched[i + 1][1]]
    up = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
    gt, eef_s, eef_e = up(gt), up(eef_s), up(eef_e)
    live_d = up(live)

    def stack(key, dtype=None):
        t = torch.stack([torch.as_tensor(ep["graph"][key]) for ep in episodes]).to(dev)
        return t if dtype is None else t.to(dtype)
    state = stack("state", torch.float32)                   # (E, n_his, N, 3)
    attrs, action = stack("attrs", torch.float32), stack("action", torch.float32)
    p_instance = stack("p_instance", torch.float32)
    state_mask, eef_mask, obj_mask = stack("state_mask"), stack("eef_mask"), stack("obj_mask")
    phys_keys = [k for k in episodes[0]["graph"].keys() if k.endswith("_physics_param")]
    phys = {k: stack(k, torch.float32) for k in phys_keys}
    N = state.shape[2]
    assert N == max_nobj + n_eef, "the graph's particle count must be max_nobj + the recording's tool points (rollout.py:105)"
    # the single-graph builder squares the threshold in Python floats (graph.py:53)
    thr2 = torch.full((E,), float(np.float32(float(adj_thresh) * float(adj_thresh))), dtype=torch.float32, device=dev)
    if all("Rr" in ep["graph"] for ep in episodes):
        edges = edges_from_onehots(stack("Rr", torch.float32), stack("Rs", torch.float32))     # truncate_graph is a no-op on CSR
    else:
        # construct_graph(dense=False) below: the start graphs' relations are the single-graph builder's on the newest history
        # frame with the same threshold (rollout/graph.py:334-337) -- built here, for all episodes in one call
        row_ptr, send, recv, n_edges, status = ops.graph_build(state[:, -1].contiguous(), state_mask, eef_mask, thr2, topk,
                                                               bool(connect_tool_all), L.AGX_SEM_SINGLE, E * max_nR)
        edges = EdgeList(row_ptr, send, recv, n_edges, status, E, N)
    n_obj = obj_mask.sum(1).clamp(min=1).to(torch.float32)
    errors = torch.zeros(E, T, dtype=torch.float32, device=dev)
    preds = torch.zeros(E, T, max_nobj, 3, dtype=torch.float32, device=dev) if return_predictions else None
    worst = edges.n_edges.max().to(torch.int32)
    overflow = edges.status.clone()

    # ---- device loop
    model.eval()
    with torch.no_grad():
        for i in range(T):
            pred_state, _ = model(state=state, attrs=attrs, p_instance=p_instance, action=action, edges=edges, **phys)
            # error = mean over the episode's object particles of |pred - gt|   (rollout.py:81-89)
            d = torch.linalg.vector_norm(pred_state - gt[:, i], dim=-1)
            errors[:, i] = (d * obj_mask).sum(1) / n_obj
            if preds is not None:
                preds[:, i] = pred_state
            if i + 1 == T:
                break
            # next graph from the prediction and the recorded tool points (rollout.py:99-137)
            states = torch.cat([pred_state, eef_s[:, i]], dim=1)                                # (E, N, 3)
            action = torch.zeros_like(states)
            action[:, max_nobj:max_nobj + n_eef] = eef_e[:, i] - eef_s[:, i]
            # episodes that have ended keep stepping on a meaningless state (tools at the origin): their particles are masked
            # out of the builder, so they own no relations and cannot raise the capacity error for the batch -- the reference
            # would have returned normally for each of them
            alive = live_d[:, i + 1]
            row_ptr, send, recv, n_edges, status = ops.graph_build(states, state_mask & alive[:, None], eef_mask, thr2, topk,
                                                                   bool(connect_tool_all), L.AGX_SEM_SINGLE, E * max_nR)
            edges = EdgeList(row_ptr, send, recv, n_edges, status, E, N)
            worst = torch.maximum(worst, n_edges.max())
            overflow |= status
            state = torch.cat([state[:, 1:], states[:, None]], dim=1)
    # ---- one synchronisation: pad_torch's failure (utils.py:37-46) and the read-back
    if int(overflow.item()) & 1 or int(worst.item()) > max_nR:
        raise RuntimeError(f"a graph reached {int(worst.item())} relations, capacity max_nR={max_nR}")
    err = errors.cpu().numpy()
    out = [[float(x) for x in err[e, :len(s)]] for e, s in enumerate(scheds)]
    if preds is not None:
        preds *= live_d[:, :, None, None]
        return out, preds
    return out


def rollout_from_start_graph(graph, fps_idx_list, dataset_config, material_config, model, device, eef_pos, obj_pos,
                             current_start, current_end, get_next_pair_or_break_func, pairs, save_dir=None, viz=False,
                             imgs=None, cam_info=None):
    """Drop-in for rollout.py:20-148 (one episode)."""
    if viz:
        raise NotImplementedError("visualisation (rollout/graph.py:visualize_graph) is outside the engine's scope; pass viz=False")
    ep = dict(graph=graph, fps_idx_list=fps_idx_list, eef_pos=eef_pos, obj_pos=obj_pos, current_start=current_start,
              current_end=current_end, pairs=pairs)
    return rollout_episodes(model, [ep], dataset_config, get_next_pair_or_break_func)[0]


def get_next_pair_or_break_episode(pairs, n_his, n_frames, current_end):
    """rollout/graph.py:374-390: the next pair starting at current_end, else walk forward to the next frame that has one."""
    valid_pairs = pairs[pairs[:, n_his - 1] == current_end]
    valid_pairs = valid_pairs[valid_pairs[:, n_his] > current_end]
    if len(valid_pairs) == 0:
        while current_end < n_frames:
            current_end += 1
            valid_pairs = pairs[pairs[:, n_his - 1] == current_end]
            valid_pairs = valid_pairs[valid_pairs[:, n_his] > current_end]
            if len(valid_pairs) > 0:
                break
        else:
            return None
    return valid_pairs[int(len(valid_pairs) / 2)]


def get_next_pair_or_break_episode_pushes(pairs, n_his, n_frames, current_end):
    """rollout/graph.py:392-400."""
    valid_pairs = pairs[pairs[:, n_his - 1] == current_end]
    valid_pairs = valid_pairs[valid_pairs[:, n_his] > current_end]
    if len(valid_pairs) == 0:
        return None
    return valid_pairs[int(len(valid_pairs) / 2)]


# ------------------------------------------------------------------------------------------ start graphs and the dataset drivers
def construct_graph(dataset_config, material_config, eef_pos, obj_pos, n_his, pair, physics_param, dense: bool = False):
    """rollout/graph.py:233-372: the start graph of one push (the reference's keys and dtypes, CPU tensors) and the kept particle
    indices.  Thinning radius and adjacency threshold are the means of their ranges; the sampler draws the reference's two numpy
    random numbers (start index among the raw particles, start index among the max_nobj survivors) and runs on the device
    (sampling.fps).  dense=False (default) leaves `Rr` / `Rs` out: `rollout_episodes` builds the start relations of all its
    episodes in one device call; dense=True adds the reference's padded one-hots."""
    from .sampling import fps
    from .utils import pad_torch
    dataset = dataset_config["datasets"][0]
    max_nobj, max_nR = dataset["max_nobj"], dataset["max_nR"]
    fps_radius = (dataset["fps_radius_range"][0] + dataset["fps_radius_range"][1]) / 2
    adj_thresh = (dataset["adj_radius_range"][0] + dataset["adj_radius_range"][1]) / 2
    n_eef, pos_dim = eef_pos.shape[1], obj_pos.shape[-1]
    N = max_nobj + n_eef
    obj_kps = np.array([obj_pos[f] for f in pair])                      # (T, N_obj_all, 3)
    eef_kps = np.array([eef_pos[f] for f in pair])                      # (T, N_eef, 3)
    fps_idx_list = fps(obj_kps[n_his - 1], max_nobj, fps_radius)
    k = len(fps_idx_list)
    state = np.zeros((n_his, N, pos_dim), np.float32)
    state[:, :k] = obj_kps[:n_his, fps_idx_list]
    state[:, max_nobj:] = eef_kps[:n_his]
    action = np.zeros((N, pos_dim), np.float32)
    action[max_nobj:] = eef_kps[n_his] - eef_kps[n_his - 1]
    state_mask, eef_mask, obj_mask = np.zeros(N, bool), np.zeros(N, bool), np.zeros(max_nobj, bool)
    state_mask[:k] = True; state_mask[max_nobj:] = True; eef_mask[max_nobj:] = True; obj_mask[:k] = True
    attrs = np.zeros((N, 2), np.float32)
    attrs[:k, 0] = 1.0; attrs[max_nobj:, 1] = 1.0
    p_instance = np.zeros((max_nobj, 1), np.float32)
    p_instance[:k, 0] = 1
    material_idx = np.zeros((max_nobj, len(material_config["material_index"])), np.int32)
    assert len(dataset_config["materials"]) == 1, "only support single material"
    material_idx[:k, material_config["material_index"][dataset_config["materials"][0]]] = 1
    graph = {
        "state": torch.from_numpy(state), "action": torch.from_numpy(action), "attrs": torch.from_numpy(attrs),
        "p_rigid": torch.zeros(1), "p_instance": torch.from_numpy(p_instance), "state_mask": torch.from_numpy(state_mask),
        "eef_mask": torch.from_numpy(eef_mask), "obj_mask": torch.from_numpy(obj_mask),
        "material_index": torch.from_numpy(material_idx).long(),
        "eef_kp": torch.from_numpy(np.stack(eef_kps[n_his - 1:n_his + 1], axis=0)).float(),
    }
    if dense:
        from .graph import construct_edges_from_states
        dev = torch.device("cuda", torch.cuda.current_device())
        Rr, Rs = construct_edges_from_states(graph["state"][-1].to(dev), adj_thresh, graph["state_mask"].to(dev), graph["eef_mask"].to(dev),
                                             dataset["topk"], dataset["connect_tool_all"])
        graph["Rr"], graph["Rs"] = pad_torch(Rr, max_nR).cpu(), pad_torch(Rs, max_nR).cpu()
    for name, v in physics_param.items():
        graph[name + "_physics_param"] = torch.from_numpy(np.asarray(v)).float()
    return graph, fps_idx_list


def _push_starts(dataset_config, material_config, eef_pos, obj_pos, episode_idx, pairs, physics_param):
    """The start graph of every push file of one episode (rollout.py:155-176), as `rollout_episodes` inputs."""
    import glob
    import os
    n_his = dataset_config["n_his"]
    pairs_path = os.path.join(dataset_config["prep_data_dir"], dataset_config["data_name"], "frame_pairs")
    eps = []
    for path in sorted(glob.glob(os.path.join(pairs_path, f"{episode_idx:06}_*.txt"))):
        valid_pairs = np.loadtxt(path).astype(int)
        pair = valid_pairs[0]              # a file holding a single row makes this a scalar and fails, as in the reference (:165-166)
        graph, fps_idx_list = construct_graph(dataset_config, material_config, eef_pos[episode_idx], obj_pos[episode_idx], n_his, pair,
                                              physics_param)
        eps.append(dict(graph=graph, fps_idx_list=fps_idx_list, eef_pos=eef_pos[episode_idx], obj_pos=obj_pos[episode_idx],
                        current_start=pair[n_his - 1], current_end=pair[n_his], pairs=pairs))
    return eps


def rollout_episode_pushes(model, device, dataset_config, material_config, eef_pos, obj_pos, episode_idx, pairs, physics_param,
                           save_dir, viz=False, imgs=None, cam_info=None):
    """rollout.py:150-203 for one episode: one error list per push file, written as `error_<i>.txt` under save_dir (no plots);
    the pushes are advanced together on the device."""
    import os
    if viz:
        raise NotImplementedError("visualisation is outside the engine's scope; pass viz=False")
    eps = _push_starts(dataset_config, material_config, eef_pos, obj_pos, episode_idx, pairs, physics_param)
    errors = rollout_episodes(model, eps, dataset_config, get_next_pair_or_break_episode_pushes) if eps else []
    for i, e in enumerate(errors):
        np.savetxt(os.path.join(save_dir, f"error_{i + 1}.txt"), np.array(e))
    return errors


def rollout_dataset(model, device, config, save_dir, viz=False):
    """rollout.py:205-268: every push of every validation episode rolled out from its start graph; writes
    `<save_dir>/<episode>/short/error_<push>.txt` and `<save_dir>/error_short.txt` ((min_steps, n_pushes) errors per step) like the
    reference, and returns that table with its per-step median and quartiles instead of drawing them.  ALL pushes of ALL
    episodes advance together in one device batch (the reference: one forward at a time, numpy round trip per step)."""
    import os
    from .dataset import load_dataset, load_positions
    if viz:
        raise NotImplementedError("visualisation is outside the engine's scope; pass viz=False")
    dataset_config, material_config = config["dataset_config"], config["material_config"]
    pair_lists, physics_params = load_dataset(dataset_config, material_config, phase="valid")
    pair_lists = np.array(pair_lists)
    eef_pos, obj_pos = load_positions(dataset_config)
    eps, owner = [], []
    for episode_idx in sorted(list(np.unique(pair_lists[:, 0]).astype(int))):
        pairs_episode = pair_lists[pair_lists[:, 0] == episode_idx][:, 1:]
        os.makedirs(os.path.join(save_dir, f"{episode_idx}", "short"), exist_ok=True)
        starts = _push_starts(dataset_config, material_config, eef_pos, obj_pos, episode_idx, pairs_episode, physics_params[episode_idx])
        owner += [(episode_idx, i + 1) for i in range(len(starts))]
        eps += starts
    total_error = rollout_episodes(model, eps, dataset_config, get_next_pair_or_break_episode_pushes)
    for (episode_idx, push), e in zip(owner, total_error):
        np.savetxt(os.path.join(save_dir, f"{episode_idx}", "short", f"error_{push}.txt"), np.array(e))
    min_step = min(len(e) for e in total_error)
    step_error = np.array([[e[i] for e in total_error] for i in range(min_step)])                 # (min_step, n_pushes)
    np.savetxt(os.path.join(save_dir, "error_short.txt"), step_error)
    return {"step_error": step_error, "median": np.median(step_error, axis=1), "p75": np.percentile(step_error, 75, axis=1),
            "p25": np.percentile(step_error, 25, axis=1), "errors": total_error, "pushes": owner}


def rollout(config, epoch, viz=False):
    """rollout.py:270-310: load `checkpoints/model_<epoch>.pth` (or `latest.pth`) written by training and evaluate it on the
    validation episodes; results under `<rollout_config.out_dir>/rollout-<data_name>-model_<epoch>`."""
    import os
    import random
    from .model import DynamicsPredictor
    dataset_config, train_config = config["dataset_config"], config["train_config"]
    seed = train_config["random_seed"]
    torch.manual_seed(seed); torch.cuda.manual_seed_all(seed); np.random.seed(seed); random.seed(seed)
    device = torch.device("cuda", torch.cuda.current_device())
    data_name = dataset_config["data_name"]
    save_dir = os.path.join(config["rollout_config"]["out_dir"], f"rollout-{data_name}-model_{epoch}")
    os.makedirs(save_dir, exist_ok=True)
    name = "latest.pth" if epoch == "latest" else f"model_{epoch}.pth"
    model = DynamicsPredictor(config["model_config"], config["material_config"], dataset_config, device)
    model.to(device)
    model.eval()
    model.load_state_dict(torch.load(os.path.join(train_config["out_dir"], data_name, "checkpoints", name), map_location=device))
    return rollout_dataset(model, device, config, save_dir, viz)
