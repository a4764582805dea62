"""Device-resident MPC rollout drivers with the reference's signatures.

Reference: src/planning/forward_dynamics.py:11-205 (`dynamics`) and :208-399 (`dynamics_masked`), action decoding
src/planning/plan_utils.py:11-20.  The reference re-enters Python for every model step (dense relation rebuild, pad,
truncate, host syncs) and builds the pusher key points on the CPU; here the per-push setup is a handful of small device
ops and the inner loop (forward_dynamics.py:156-197 / :351-393) is a handful of `DynamicsPredictor.rollout` calls: samples are
sorted by their repeat count and leave the batch once their prediction is captured (`action_repeat == ai`, :160-161), so a push
costs sum(repeat) model steps instead of bsz * max(repeat).
"""
from __future__ import annotations

import torch


def decode_action(action, push_length=0.10):
    """plan_utils.py:11-20: (x, z, theta, length) -> (x_start, z_start, x_end, z_end), repeat count = int(length)."""
    x_start, z_start, theta = action[:, :, 0], action[:, :, 1], action[:, :, 2]
    action_repeat = action[:, :, 3].detach().to(torch.int32)
    x_end = x_start - push_length * torch.cos(theta)
    z_end = z_start - push_length * torch.sin(theta)
    return torch.stack([x_start, z_start, x_end, z_end], dim=-1), action_repeat


def _pusher_keypoints(task_config, decoded, theta, y, sim_real_ratio):
    """End-effector key points and per-step deltas of one push (forward_dynamics.py:42-81); decoded (bsz, 4), y (bsz,)."""
    bsz, dev = decoded.shape[0], decoded.device
    pts = task_config["pusher_points"]
    dx = decoded[:, 2] - decoded[:, 0]
    dz = decoded[:, 3] - decoded[:, 1]
    if len(pts) == 1:
        eef = torch.stack([decoded[:, 0], y, decoded[:, 1]], -1)[:, None]
        delta = torch.stack([dx, torch.zeros_like(dx), dz], -1)[:, None]
    elif len(pts) == 5:
        off = torch.tensor([0.0] + [float(pts[i][1]) * sim_real_ratio for i in range(1, 5)], device=dev, dtype=decoded.dtype)
        ex = decoded[:, 0:1] + off[None] * torch.sin(theta)[:, None]
        ez = decoded[:, 1:2] - off[None] * torch.cos(theta)[:, None]
        eef = torch.stack([ex, y[:, None].expand(bsz, 5), ez], -1)
        delta = torch.stack([dx, torch.zeros_like(dx), dz], -1)[:, None].expand(bsz, 5, 3).contiguous()
    else:
        raise NotImplementedError("pusher not implemented")
    if task_config["gripper_enable"]:
        eef = eef.clone()
        eef[:, :, 1] += 0.01 * sim_real_ratio
    return eef, delta


def _physics(ppm_optimizer, physics_param, bsz, dev):
    physics_param = ppm_optimizer.physics_param if physics_param is None else physics_param
    name = ppm_optimizer.material
    if name in physics_param:
        return physics_param[name].to(dev, torch.float32)[None].repeat(bsz, 1)
    return torch.zeros((bsz, ppm_optimizer.material_dims[name]), dtype=torch.float32, device=dev)


def _rollout_captured(model, states, attrs, delta, p_instance, phys, mask, eef_mask, adj_thresh, topk, cta, repeat, max_nR, y_mode, raise_):
    """Prediction of every sample after ITS OWN number of model steps (forward_dynamics.py:160-161 captures `action_repeat == ai`
    inside a loop that runs every sample to the largest count); zeros where repeat < 1.  Samples are independent graphs, so they are
    sorted by repeat count and leave the batch as soon as they are captured: the work is sum(repeat) model steps instead of
    bsz * max(repeat) (1.8x less for counts spread uniformly over 1..8).  One host read (the sorted counts; the reference reads
    the maximum, :156) and one capacity check at the end."""
    bsz, n_obj, dev = states.shape[0], p_instance.shape[1], states.device
    pred = torch.zeros((bsz, n_obj, 3), device=dev)
    rep_sorted, order = torch.sort(repeat.long(), descending=True, stable=True)
    counts = rep_sorted.tolist()
    active = sum(1 for c in counts if c >= 1)
    if active == 0:
        return pred
    take = lambda t: t[order[:active]].contiguous()  # noqa: E731
    hist, attrs, delta, p_instance, phys, mask, eef_mask = (take(t) for t in (states, attrs, delta, p_instance, phys, mask, eef_mask))
    per_sample_thr = torch.is_tensor(adj_thresh) and adj_thresh.numel() == bsz and bsz > 1
    thr = adj_thresh.reshape(-1)[order[:active]] if per_sample_thr else adj_thresh
    worst = torch.zeros((), dtype=torch.int32, device=dev)
    overflow = torch.zeros(1, dtype=torch.int32, device=dev)
    done = 0
    while active > 0:
        r = counts[active - 1]                                   # smallest remaining count: its samples are captured by this segment
        first = active - 1
        while first > 0 and counts[first - 1] == r:
            first -= 1
        o = model.rollout(hist[:active], attrs[:active], delta[:active], p_instance[:active], phys[:active], mask[:active], eef_mask[:active],
                          thr[:active] if per_sample_thr else thr, topk, cta, r - done, max_nR,
                          y_mode=y_mode, gripper_raise=raise_, check=False)
        pred[order[first:active]] = o["state_seqs"][first:active, -1]
        worst = torch.maximum(worst, o["n_edges"].max())
        overflow |= o["status"]
        hist = o["state"]
        done, active = r, first
    if int(overflow.item()) & 1 or int(worst.item()) > max_nR:
        raise RuntimeError(f"rollout: a graph reached {int(worst.item())} relations, capacity max_nR={max_nR}")
    return pred


@torch.no_grad()
def dynamics(state, action, model, device, ppm_optimizer, physics_param=None):
    """Drop-in for forward_dynamics.py:11-205.  state (n_obj, 3), action (bsz, n_look_forward, 4)."""
    tc = ppm_optimizer.task_config
    max_nR, n_his, ratio = tc["max_nR"], tc["n_his"], tc["sim_real_ratio"]
    state, action = state.to(device), action.to(device)
    bsz, n_look = action.shape[0], action.shape[1]
    decoded, repeat = decode_action(action, push_length=tc["push_length"])
    n_obj, n_eef = state.shape[0], ppm_optimizer.eef_num
    N = n_obj + n_eef
    raise_ = 0.01 * ratio if tc["gripper_enable"] else 0.0
    pred_seq = torch.zeros((bsz, n_look, n_obj, 3), device=device)
    attrs = torch.zeros((bsz, N, 2), device=device)
    attrs[:, :n_obj, 0] = 1.0
    attrs[:, n_obj:, 1] = 1.0
    p_instance = torch.ones((bsz, n_obj, 1), device=device)
    state_mask = torch.ones((bsz, N), dtype=torch.bool, device=device)
    eef_mask = torch.zeros((bsz, N), dtype=torch.bool, device=device)
    eef_mask[:, n_obj:] = True
    phys = _physics(ppm_optimizer, physics_param, bsz, device)
    for li in range(n_look):
        obj = state[None].repeat(bsz, 1, 1) if li == 0 else pred_seq[:, li - 1]
        y = obj[:, :, 1].min(dim=1).values
        eef, delta = _pusher_keypoints(tc, decoded[:, li], action[:, li, 2], y, ratio)
        cur = torch.cat([obj, eef], 1)
        states = cur[:, None].repeat(1, n_his, 1, 1)
        states_delta = torch.zeros((bsz, N, 3), device=device)
        states_delta[:, n_obj:] = delta
        pred_seq[:, li] = _rollout_captured(model, states, attrs, states_delta, p_instance, phys, state_mask, eef_mask, ppm_optimizer.adj_thresh,
                                            tc["topk"], tc["connect_tools_all"], repeat[:, li], max_nR, "min", raise_)
    return {"state_seqs": pred_seq, "action_seqs": decoded}


@torch.no_grad()
def dynamics_masked(state_init, state_mask, action, model, device, ppm_optimizer, physics_param=None):
    """Drop-in for forward_dynamics.py:208-399.  state_init (bsz, n_obj, 3), state_mask (bsz, n_obj) bool, action (bsz, 4)."""
    tc = ppm_optimizer.task_config
    max_nR, n_his, ratio = tc["max_nR"], tc["n_his"], tc["sim_real_ratio"]
    state, state_mask, action = state_init.to(device), state_mask.to(device), action.to(device)
    bsz, n_obj = state.shape[0], state.shape[1]
    decoded, repeat = decode_action(action[:, None], push_length=tc["push_length"])
    decoded, repeat = decoded[:, 0], repeat[:, 0]
    n_eef = ppm_optimizer.eef_num
    N = n_obj + n_eef
    raise_ = 0.01 * ratio if tc["gripper_enable"] else 0.0
    m = state_mask.to(state.dtype)
    y = (state[:, :, 1] * m).sum(dim=1) / m.sum(dim=1)
    eef, delta = _pusher_keypoints(tc, decoded, action[:, 2], y, ratio)
    cur = torch.cat([state, eef], 1)
    states = cur[:, None].repeat(1, n_his, 1, 1)
    states_delta = torch.zeros((bsz, N, 3), device=device)
    states_delta[:, n_obj:] = delta
    attrs = torch.zeros((bsz, N, 2), device=device)
    attrs[:, :n_obj, 0] = m
    attrs[:, n_obj:, 1] = 1.0
    counts = state_mask.sum(1)
    p_instance = (torch.arange(n_obj, device=device)[None] < counts[:, None]).to(state.dtype)[:, :, None]   # :304-310 prefix fill
    mask_new = torch.cat([state_mask, torch.ones((bsz, n_eef), dtype=torch.bool, device=device)], 1)
    eef_mask = torch.zeros((bsz, N), dtype=torch.bool, device=device)
    eef_mask[:, n_obj:] = True
    phys = _physics(ppm_optimizer, physics_param, bsz, device)
    pred = _rollout_captured(model, states, attrs, states_delta, p_instance, phys, mask_new, eef_mask, ppm_optimizer.adj_thresh, tc["topk"],
                             tc["connect_tools_all"], repeat, max_nR, "masked_mean", raise_)
    return {"state_seqs": pred, "action_seqs": decoded}
