"""Batched evaluation rollout, replacing the batch-1, numpy-round-trip loop of
src/dynamics/rollout/rollout.py:20-148.

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
    edges = edges_from_onehots(stack("Rr", torch.float32), stack("Rs", torch.float32))         # truncate_graph is a no-op on CSR
    # the single-graph builder squares the threshold in Python floats (graph.py:53)
    thr2 = torch.full((E,), float(np.float32(float(adj_thresh) * float(adj_thresh))), dtype=torch.float32, device=dev)
    n_obj = obj_mask.sum(1).clamp(min=1).to(torch.float32)
    errors = torch.zeros(E, T, dtype=torch.float32, device=dev)
    preds = torch.zeros(E, T, max_nobj, 3, dtype=torch.float32, device=dev) if return_predictions else None
    worst = torch.zeros((), dtype=torch.int32, device=dev)
    overflow = torch.zeros(1, dtype=torch.int32, device=dev)

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
