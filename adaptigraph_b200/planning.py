"""Device-resident MPC rollout drivers with the reference's signatures.

Reference: src/planning/forward_dynamics.py:11-205 (`dynamics`) and :208-399 (`dynamics_masked`), action decoding
src/planning/plan_utils.py:11-20.  The reference re-enters Python for every model step (dense relation rebuild, pad,
truncate, host syncs) and builds the pusher key points on the CPU; here the per-push setup is a handful of small device
ops and the whole inner loop (forward_dynamics.py:156-197 / :351-393) is ONE `DynamicsPredictor.rollout` call.  Each
sample's prediction is captured at its own repeat count (`action_repeat == ai`, :160-161) by indexing the rollout output.
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


def _capture(seq, repeat):
    """seq (bsz, T, n, 3), repeat (bsz,) int: prediction after `repeat` model steps; zeros where repeat < 1 (never captured)."""
    idx = (repeat.long() - 1).clamp(min=0)
    out = seq[torch.arange(seq.shape[0], device=seq.device), idx]
    return torch.where((repeat >= 1)[:, None, None], out, torch.zeros_like(out))


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
        T = int(repeat[:, li].max().item())              # the reference syncs here too (:156)
        if T < 1:
            continue
        out = model.rollout(states, attrs, states_delta, p_instance, phys, state_mask, eef_mask, ppm_optimizer.adj_thresh,
                            tc["topk"], tc["connect_tools_all"], T, max_nR, y_mode="min", gripper_raise=raise_)
        pred_seq[:, li] = _capture(out["state_seqs"], repeat[:, li])
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
    pred = torch.zeros((bsz, n_obj, 3), device=device)
    T = int(repeat.max().item())
    if T >= 1:
        out = model.rollout(states, attrs, states_delta, p_instance, phys, mask_new, eef_mask, ppm_optimizer.adj_thresh, tc["topk"],
                            tc["connect_tools_all"], T, max_nR, y_mode="masked_mean", gripper_raise=raise_)
        pred = _capture(out["state_seqs"], repeat)
    return {"state_seqs": pred, "action_seqs": decoded}
