"""MPC reward terms evaluated on the device after the rollout, with the reference's signatures.

    chamfer(x, y)                                       src/planning/losses.py:4-10
    box_loss(state, target)                             losses.py:25-35
    rope_penalty / cloth_penalty / granular_penalty     losses.py:37-92
    running_cost(state, action, state_cur, error_func, penalty_func, bbox)               src/planning/plan.py:27-59

This module holds NO arithmetic: `chamfer` is the stand-alone kernel `agx_chamfer`; `running_cost` — what the planner actually
calls (plan.py:175 binds it with `partial(chamfer, y=...)` / `partial(box_loss, target=...)` and one of the three penalties) —
recognises those bound callables and runs the whole tail as ONE kernel (`agx_running_cost`, csrc/rewards.cu): error term, collision
penalty, workspace-box penalty, the two batch-wide maxima and the final weighting, with no host synchronisation (the reference
synchronises twice, for `error.max().item()` and cloth's `max().item()`, and once more for its progress print).  The stand-alone
`box_loss` and `*_penalty` functions evaluate the same kernel on the matching slice of its work and read the term back from it.
"""
from __future__ import annotations

import functools

import torch

from . import _lib as L
from . import ops


def _on_device(*ts):
    for t in ts:
        if not t.is_cuda:
            raise RuntimeError("adaptigraph_b200.rewards evaluates on CUDA tensors only (no CPU fallback); got a "
                               f"{t.device} tensor of shape {tuple(t.shape)}")


def chamfer(x, y):  # x: (B, N, D), y: (1|B, M, D)
    return ops.chamfer(x, y)


def _terms(state, action, state_cur, bbox, target, error_mode, penalty_mode, sim_real_ratio):
    """(bsz, L, 4) per-cell terms [error, penalty, -, box penalty] of the fused kernel's workspace after a call."""
    _on_device(state, action, state_cur)
    ops.running_cost(state, action, state_cur, bbox, target, error_mode, penalty_mode, sim_real_ratio)
    bsz, Lk = state.shape[:2]
    ws = ops._RC_WS[(state.device.index, int(L.lib.agx_running_cost_workspace_bytes(bsz, Lk)))]
    return ws[: bsz * Lk * 16].view(torch.float32).view(bsz, Lk, 4)


def _unit_box(dev):
    return torch.tensor([[-1e30, 1e30], [-1e30, 1e30]], dtype=torch.float32, device=dev)


def box_loss(state, target):
    """losses.py:25-35: state (B, N, 3), target (2, 2) -> (B)."""
    _on_device(state, target)
    B, n, _ = state.shape
    act = torch.zeros(B, 1, 3, dtype=torch.float32, device=state.device)
    t = _terms(state.reshape(B, 1, n, 3), act, state[0], _unit_box(state.device), target, L.AGX_ERROR_BOX, L.AGX_PENALTY_ROPE, 10.0)
    return t[:, 0, 0].clone()


def _penalty(mode, state_pred, action, state_init, sim_real_ratio):
    _on_device(state_pred, action, state_init)
    dev = state_pred.device
    t = _terms(state_pred, action, state_init, _unit_box(dev), _unit_box(dev), L.AGX_ERROR_BOX, mode, sim_real_ratio)
    return t[:, :, 1].clone()


def rope_penalty(state_pred, action, state_init, sim_real_ratio=10.0):
    return _penalty(L.AGX_PENALTY_ROPE, state_pred, action, state_init, sim_real_ratio)


def cloth_penalty(state_pred, action, state_init, sim_real_ratio=10.0):
    return _penalty(L.AGX_PENALTY_CLOTH, state_pred, action, state_init, sim_real_ratio)


def granular_penalty(state_pred, action, state_init, sim_real_ratio=10.0):
    return _penalty(L.AGX_PENALTY_GRANULAR, state_pred, action, state_init, sim_real_ratio)


_PENALTIES = {"rope_penalty": L.AGX_PENALTY_ROPE, "cloth_penalty": L.AGX_PENALTY_CLOTH, "granular_penalty": L.AGX_PENALTY_GRANULAR}


def _unwrap(fn):
    """(function, keyword arguments) of a possibly functools.partial-bound callable (plan.py:146, :155, :160-165)."""
    kw = {}
    while isinstance(fn, functools.partial):
        if fn.args:
            raise NotImplementedError("running_cost: positional arguments bound into error_func / penalty_func are not supported")
        kw = {**fn.keywords, **kw}
        fn = fn.func
    return fn, kw


def running_cost(state, action, state_cur, error_func, penalty_func, bbox, verbose=False, **kwargs):
    """plan.py:27-59.  state (bsz, n_look_forward, max_nobj, 3), action (bsz, n_look_forward, action_dim), state_cur
    (max_nobj, 3), bbox (2, 2) -> {'reward_seqs': (bsz,)}.  error_func / penalty_func are the planner's bound callables
    (`partial(chamfer, y=target[None])` or `partial(box_loss, target=box)`; `partial(<material>_penalty, sim_real_ratio=r)`),
    from this module or the reference's planning.losses — they are recognised by name, not called."""
    _on_device(state, action, state_cur)
    efn, ekw = _unwrap(error_func)
    pfn, pkw = _unwrap(penalty_func)
    ename, pname = getattr(efn, "__name__", ""), getattr(pfn, "__name__", "")
    if ename == "chamfer" and "y" in ekw:
        y = torch.as_tensor(ekw["y"])
        if y.dim() == 3:
            if y.shape[0] != 1:
                raise NotImplementedError("running_cost: the fused kernel takes one target point set shared by all samples (plan.py:146)")
            y = y[0]
        mode, target = L.AGX_ERROR_CHAMFER, y
    elif ename == "box_loss" and "target" in ekw:
        mode, target = L.AGX_ERROR_BOX, torch.as_tensor(ekw["target"])
    else:
        raise NotImplementedError(f"running_cost: error_func must be chamfer bound with y= or box_loss bound with target= (plan.py:146, :155); got {error_func!r}")
    if pname not in _PENALTIES:
        raise NotImplementedError(f"running_cost: penalty_func must be rope_penalty / cloth_penalty / granular_penalty (plan.py:160-165); got {penalty_func!r}")
    dev = state.device
    bbox = torch.as_tensor(bbox, dtype=torch.float32).to(dev)               # plan.py builds it as a numpy array
    reward = ops.running_cost(state, action, state_cur, bbox, target.to(dev, torch.float32), mode, _PENALTIES[pname],
                              float(pkw.get("sim_real_ratio", 10.0)))
    if verbose:                                                              # the reference's progress print (two host syncs)
        print(f"max reward {reward.max().item()}")
    return {"reward_seqs": reward}
