"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the MPC rollout drivers
src/planning/forward_dynamics.py:11-205 (`dynamics`) and :208-399 (`dynamics_masked`), on top of
oracle.dynamics_oracle.rollout_dense.  Pinned by tests/test_oracle_golden.py against
tests/golden/planning_dynamics.npz (outputs of the reference's own functions)."""
import math

import torch

from . import dynamics_oracle as orc


def decode(action, push_length):
    """plan_utils.py:11-20"""
    x, z, th = action[..., 0], action[..., 1], action[..., 2]
    return torch.stack([x, z, x - push_length * torch.cos(th), z - push_length * torch.sin(th)], -1), action[..., 3].to(torch.int32)


def _eef(pusher, ratio, dec, theta, y, gripper):
    bsz = dec.shape[0]
    k = len(pusher)
    eef = torch.zeros(bsz, k, 3)
    delta = torch.zeros(bsz, k, 3)
    delta[:, :, 0] = (dec[:, 2] - dec[:, 0])[:, None]
    delta[:, :, 2] = (dec[:, 3] - dec[:, 1])[:, None]
    eef[:, :, 1] = y[:, None]
    for i in range(k):
        off = float(pusher[i][1]) * ratio if (k == 5 and i > 0) else 0.0          # forward_dynamics.py:64-75
        eef[:, i, 0] = dec[:, 0] + off * torch.sin(theta)
        eef[:, i, 2] = dec[:, 1] - off * torch.cos(theta)
    if gripper:
        eef[:, :, 1] += 0.01 * ratio                                                # :80-81
    return eef, delta


def _step_capture(p, pstep, states, attrs, p_inst, delta_full, phys, mask, eef_mask, thr, topk, cta, repeat, y_mode, raise_):
    T = int(repeat.max())
    out = torch.zeros(states.shape[0], p_inst.shape[1], 3)
    if T < 1:
        return out
    preds, _ = orc.rollout_dense(p, pstep, states, attrs, p_inst, delta_full, phys, mask, eef_mask, thr, topk, cta, T,
                                 y_mode=y_mode, gripper_raise=raise_)
    for b in range(states.shape[0]):
        if int(repeat[b]) >= 1:
            out[b] = preds[b, int(repeat[b]) - 1]                                   # :160-161
    return out


def dynamics(p, pstep, state, action, cfg):
    """cfg: dict(pusher, ratio, push_length, gripper, thr, topk, cta, n_his, phys)"""
    bsz, n_look = action.shape[:2]
    dec, rep = decode(action, cfg["push_length"])
    n_obj, k = state.shape[0], len(cfg["pusher"])
    N = n_obj + k
    seq = torch.zeros(bsz, n_look, n_obj, 3)
    attrs = torch.zeros(bsz, N, 2); attrs[:, :n_obj, 0] = 1; attrs[:, n_obj:, 1] = 1
    mask = torch.ones(bsz, N, dtype=torch.bool)
    eef_mask = torch.zeros(bsz, N, dtype=torch.bool); eef_mask[:, n_obj:] = True
    p_inst = torch.ones(bsz, n_obj, 1)
    phys = torch.full((bsz, 1), cfg["phys"])
    raise_ = 0.01 * cfg["ratio"] if cfg["gripper"] else 0.0
    for li in range(n_look):
        obj = state[None].repeat(bsz, 1, 1) if li == 0 else seq[:, li - 1]
        y = obj[:, :, 1].min(1).values
        eef, delta = _eef(cfg["pusher"], cfg["ratio"], dec[:, li], action[:, li, 2], y, cfg["gripper"])
        cur = torch.cat([obj, eef], 1)
        states = cur[:, None].repeat(1, cfg["n_his"], 1, 1)
        dfull = torch.zeros(bsz, N, 3); dfull[:, n_obj:] = delta
        seq[:, li] = _step_capture(p, pstep, states, attrs, p_inst, dfull, phys, mask, eef_mask, cfg["thr"], cfg["topk"], cfg["cta"],
                                   rep[:, li], "min", raise_)
    return seq, dec


def dynamics_masked(p, pstep, state, state_mask, action, cfg):
    bsz, n_obj = state.shape[:2]
    dec, rep = decode(action[:, None], cfg["push_length"])
    dec, rep = dec[:, 0], rep[:, 0]
    k = len(cfg["pusher"])
    N = n_obj + k
    m = state_mask.float()
    y = (state[:, :, 1] * m).sum(1) / m.sum(1)
    eef, delta = _eef(cfg["pusher"], cfg["ratio"], dec, action[:, 2], y, cfg["gripper"])
    cur = torch.cat([state, eef], 1)
    states = cur[:, None].repeat(1, cfg["n_his"], 1, 1)
    dfull = torch.zeros(bsz, N, 3); dfull[:, n_obj:] = delta
    attrs = torch.zeros(bsz, N, 2); attrs[:, :n_obj, 0] = m; attrs[:, n_obj:, 1] = 1
    p_inst = torch.zeros(bsz, n_obj, 1)
    for b in range(bsz):
        p_inst[b, :int(state_mask[b].sum()), 0] = 1                                 # :304-310 (prefix, not the mask positions)
    mask = torch.cat([state_mask, torch.ones(bsz, k, dtype=torch.bool)], 1)
    eef_mask = torch.zeros(bsz, N, dtype=torch.bool); eef_mask[:, n_obj:] = True
    phys = torch.full((bsz, 1), cfg["phys"])
    raise_ = 0.01 * cfg["ratio"] if cfg["gripper"] else 0.0
    return _step_capture(p, pstep, states, attrs, p_inst, dfull, phys, mask, eef_mask, cfg["thr"], cfg["topk"], cfg["cta"], rep,
                         "masked_mean", raise_), dec


# ---------------------------------------------------------------------------------------------- reward terms
def chamfer(x, y):
    """planning/losses.py:4-10, line by line (CPU torch).  Pinned: tests/golden/rewards.npz holds the reference's outputs."""
    x = x[:, None].repeat(1, y.shape[1], 1, 1)
    y = y[:, :, None].repeat(1, 1, x.shape[2], 1)
    dis = torch.norm(x - y, 2, dim=-1)
    return torch.mean(dis.min(dim=2).values, dim=1) + torch.mean(dis.min(dim=1).values, dim=1)


def running_cost(state, action, state_cur, error_func, penalty_func, bbox):
    """planning/plan.py:27-59 restated (plan.py itself cannot be imported here: pyflex / GroundingDINO / SAM are absent).
    PINNED: tests/golden/running_cost.npz holds outputs of the reference's own function, whose unmodified source is executed by
    tests/golden/make_golden_running_cost.py; every term it calls is pinned through tests/golden/rewards.npz."""
    bsz, n_look_forward = state.shape[0], state.shape[1]
    state_flat = state.reshape(bsz * n_look_forward, state.shape[2], state.shape[3])
    error = error_func(state_flat).reshape(bsz, n_look_forward)
    error_weight = 2. / (error.max().item() + 1e-6)
    collision_penalty = penalty_func(state, action, state_cur)
    xmax = state.max(dim=2).values[:, :, 0]
    xmin = state.min(dim=2).values[:, :, 0]
    zmax = state.max(dim=2).values[:, :, 2]
    zmin = state.min(dim=2).values[:, :, 2]
    box_penalty = torch.stack([
        torch.maximum(xmin - bbox[0, 0], torch.zeros_like(xmin)),
        torch.maximum(bbox[0, 1] - xmax, torch.zeros_like(xmax)),
        torch.maximum(zmin - bbox[1, 0], torch.zeros_like(zmin)),
        torch.maximum(bbox[1, 1] - zmax, torch.zeros_like(zmax)),
    ], dim=-1)
    box_penalty = torch.exp(-box_penalty * 100.).max(dim=-1).values
    return -error_weight * error[:, -1] - 5. * collision_penalty.mean(dim=1) - 5. * box_penalty.mean(dim=1)


def rope_penalty(state_pred, action, state_init, sim_real_ratio=10.0):
    """planning/losses.py:37-48, line by line (CPU torch).  Pinned: tests/golden/rewards.npz holds the reference's output."""
    bsz, n_look_forward, _ = action.shape
    x_start = action[:, :, 0]
    z_start = action[:, :, 1]
    action_point_2d = torch.stack([x_start, z_start], dim=-1)
    state_2d = torch.cat([state_init[:, [0, 2]][None, None].repeat(bsz, 1, 1, 1), state_pred[:, :-1, :, [0, 2]]], dim=1)
    action_state_distance = torch.norm(action_point_2d[:, :, None] - state_2d, dim=-1).min(dim=-1).values
    pusher_size = 0.02 * sim_real_ratio
    action_state_distance = torch.maximum(action_state_distance - pusher_size, torch.zeros_like(action_state_distance))
    return torch.exp(-action_state_distance * 100.)


def cloth_penalty(state_pred, action, state_init, sim_real_ratio=10.0):
    """planning/losses.py:51-65 (CPU torch).  Pinned: tests/golden/rewards.npz holds the reference's output."""
    action_point_2d = torch.stack([action[:, :, 0], action[:, :, 1]], dim=-1)
    dist = torch.norm(action_point_2d[:, :, None] - state_init[:, [0, 2]][None, None], dim=-1)
    dmin = dist.min(dim=-1).values
    dmin = torch.maximum(dmin - 0.005 * sim_real_ratio, torch.zeros_like(dmin))
    dmax = torch.minimum(dist.max(dim=-1).values, torch.ones_like(dmin) * 0.4 * sim_real_ratio)
    dmax = dmax / dmax.max().item()
    return 1. - torch.exp(-dmin * 100.) - dmax * 0.2


def granular_penalty(state_pred, action, state_init, sim_real_ratio=10.0):
    """planning/losses.py:67-92 (CPU torch).  Pinned: tests/golden/rewards.npz holds the reference's output."""
    bsz, n_look_forward, _ = action.shape
    x_start, z_start, theta = action[:, :, 0], action[:, :, 1], action[:, :, 2]
    pusher_radius = 0.05 * sim_real_ratio
    delta_x = pusher_radius * torch.sin(theta)
    delta_z = -pusher_radius * torch.cos(theta)
    pts = []
    for f in (1., 0.75, 0.5, 0.25):
        pts += [x_start - delta_x if f == 1. else x_start - f * delta_x, z_start - delta_z if f == 1. else z_start - f * delta_z]
    pts += [x_start, z_start]
    for f in (0.25, 0.5, 0.75, 1.):
        pts += [x_start + delta_x if f == 1. else x_start + f * delta_x, z_start + delta_z if f == 1. else z_start + f * delta_z]
    action_point_2d = torch.stack(pts, dim=-1).reshape(bsz, n_look_forward, 9, 2)
    state_2d = torch.cat([state_init[:, [0, 2]][None, None].repeat(bsz, 1, 1, 1), state_pred[:, :-1, :, [0, 2]]], dim=1)
    d = torch.norm(action_point_2d[:, :, :, None] - state_2d[:, :, None], dim=-1).min(dim=-1).values.min(dim=-1).values
    d = torch.maximum(d - 0.02 * sim_real_ratio, torch.zeros_like(d))
    return torch.exp(-d * 100.)
