"""MPC reward terms evaluated on the device after the rollout, with the reference's signatures.

    chamfer(x, y)                                       src/planning/losses.py:4-10     -> CUDA kernel (agx_chamfer)
    box_loss(state, target)                             losses.py:25-35
    rope_penalty / cloth_penalty / granular_penalty     losses.py:37-92
    running_cost(state, action, state_cur, error_func, penalty_func, bbox)               src/planning/plan.py:27-59

`chamfer` is the term that costs: the reference materialises two (B, M, N, 3) tensors per call (B = samples x look-ahead
steps); here it is one kernel that never leaves shared memory.  The remaining terms are a few element-wise device ops on
(bsz, n_look_forward, max_nobj) tensors and are restated with torch device ops unchanged in arithmetic.  The reference's
`running_cost` synchronises for its progress print and for `error.max().item()`; here the weight stays on the device and the
print is opt-in, so planning needs no host round trip between rollout and reward.
"""
from __future__ import annotations

import torch

from . import ops


def _on_device(*ts):
    for t in ts:
        if not t.is_cuda:
            raise RuntimeError("adaptigraph_b200.rewards evaluates on CUDA tensors only (no CPU fallback); got a "
                               f"{t.device} tensor of shape {tuple(t.shape)}")


def chamfer(x, y):  # x: (B, N, D), y: (1|B, M, D)
    return ops.chamfer(x, y)


def box_loss(state, target):
    _on_device(state, target)
    xmin, xmax, zmin, zmax = target[0, 0], target[0, 1], target[1, 0], target[1, 1]
    zero = torch.zeros_like(state[:, :, 0])
    x_diff = torch.maximum(xmin - state[:, :, 0], zero) + torch.maximum(state[:, :, 0] - xmax, zero)
    z_diff = torch.maximum(zmin - state[:, :, 2], zero) + torch.maximum(state[:, :, 2] - zmax, zero)
    return ((x_diff ** 2 + z_diff ** 2) ** 0.5).mean(dim=1)


def _min_point_distance(action_point_2d, state_2d):
    """min over particles of |point - particle| for (bsz, L, 2) points against (bsz|1, L|1, n, 2) particles."""
    return torch.norm(action_point_2d[:, :, None] - state_2d, dim=-1).min(dim=-1).values


def rope_penalty(state_pred, action, state_init, sim_real_ratio=10.0):
    _on_device(state_pred, action, state_init)
    bsz = action.shape[0]
    action_point_2d = torch.stack([action[:, :, 0], action[:, :, 1]], dim=-1)
    state_2d = torch.cat([state_init[:, [0, 2]][None, None].expand(bsz, 1, -1, -1), state_pred[:, :-1, :, [0, 2]]], dim=1)
    d = _min_point_distance(action_point_2d, state_2d)
    d = torch.maximum(d - 0.02 * sim_real_ratio, torch.zeros_like(d))
    return torch.exp(-d * 100.)


def cloth_penalty(state_pred, action, state_init, sim_real_ratio=10.0):
    _on_device(state_pred, action, state_init)
    action_point_2d = torch.stack([action[:, :, 0], action[:, :, 1]], dim=-1)
    dist = torch.norm(action_point_2d[:, :, None] - state_init[:, [0, 2]][None, None], dim=-1)
    dmin = dist.min(dim=-1).values
    dmin = torch.maximum(dmin - 0.005 * sim_real_ratio, torch.zeros_like(dmin))
    dmax = torch.minimum(dist.max(dim=-1).values, torch.full_like(dmin, 0.4 * sim_real_ratio))
    dmax = dmax / dmax.max()                                           # losses.py:63 (.item() there: a host sync; not here)
    return 1. - torch.exp(-dmin * 100.) - dmax * 0.2


def granular_penalty(state_pred, action, state_init, sim_real_ratio=10.0):
    _on_device(state_pred, action, state_init)
    bsz, n_look_forward, _ = action.shape
    x_start, z_start, theta = action[:, :, 0], action[:, :, 1], action[:, :, 2]
    pusher_radius = 0.05 * sim_real_ratio
    delta_x = pusher_radius * torch.sin(theta)
    delta_z = -pusher_radius * torch.cos(theta)
    pts = []
    for f in (-1., -0.75, -0.5, -0.25, 0., 0.25, 0.5, 0.75, 1.):     # the 9 points along the pusher blade, losses.py:75-84
        pts += [x_start + f * delta_x if f else x_start, z_start + f * delta_z if f else z_start]
    action_point_2d = torch.stack(pts, dim=-1).reshape(bsz, n_look_forward, 9, 2)
    state_2d = torch.cat([state_init[:, [0, 2]][None, None].expand(bsz, 1, -1, -1), state_pred[:, :-1, :, [0, 2]]], dim=1)
    d = torch.norm(action_point_2d[:, :, :, None] - state_2d[:, :, None], dim=-1).min(dim=-1).values.min(dim=-1).values
    d = torch.maximum(d - 0.02 * sim_real_ratio, torch.zeros_like(d))
    return torch.exp(-d * 100.)


def running_cost(state, action, state_cur, error_func, penalty_func, bbox, verbose=False, **kwargs):
    """plan.py:27-59.  state (bsz, n_look_forward, max_nobj, 3), action (bsz, n_look_forward, action_dim), state_cur
    (max_nobj, 3), bbox (2, 2) -> {'reward_seqs': (bsz,)}."""
    _on_device(state, action, state_cur, bbox)
    bsz, n_look_forward = state.shape[0], state.shape[1]
    state_flat = state.reshape(bsz * n_look_forward, state.shape[2], state.shape[3])
    error = error_func(state_flat).reshape(bsz, n_look_forward)
    error_weight = 2. / (error.max() + 1e-6)
    collision_penalty = penalty_func(state, action, state_cur)
    mx, mn = state.max(dim=2).values, state.min(dim=2).values
    xmax, xmin, zmax, zmin = mx[:, :, 0], mn[:, :, 0], mx[:, :, 2], mn[:, :, 2]
    zero = torch.zeros_like(xmin)
    box_penalty = torch.stack([torch.maximum(xmin - bbox[0, 0], zero), torch.maximum(bbox[0, 1] - xmax, zero),
                               torch.maximum(zmin - bbox[1, 0], zero), torch.maximum(bbox[1, 1] - zmax, zero)], dim=-1)
    box_penalty = torch.exp(-box_penalty * 100.).max(dim=-1).values
    reward = -error_weight * error[:, -1] - 5. * collision_penalty.mean(dim=1) - 5. * box_penalty.mean(dim=1)
    if verbose:
        print(f'min error {error[:, -1].min().item()}, max reward {reward.max().item()}')
    return {"reward_seqs": reward}
