"""Autograd bridge for DynamicsPredictor.forward: the training unroll of train.py:90-112 differentiates through the
model (parameters and, for BPTT, `state`).  Forward = agx_forward_train (exact fp32, activations saved in one opaque
buffer), backward = agx_backward (deterministic reductions).  No eager fallback."""
from __future__ import annotations

import torch

from . import ops


def _sender_lists(row_ptr, send, recv, B, N):
    """The same relations grouped by flattened sender id, in a fixed (stable) order: send_ptr (B*N+1), send_perm (E_cap).
    Device-side torch plumbing (stable sort + integer index_add), no host synchronisation, so it can sit inside a captured
    CUDA graph."""
    E_cap = send.numel()
    e = torch.arange(E_cap, device=send.device)
    valid = e < row_ptr[-1]
    key = torch.where(valid, (recv.long() // N) * N + send.long(), torch.full_like(e, B * N))
    key_sorted, perm = torch.sort(key, stable=True)
    deg = torch.zeros(B * N + 1, dtype=torch.int64, device=send.device).index_add_(0, key_sorted, torch.ones_like(key_sorted))[: B * N]
    send_ptr = torch.zeros(B * N + 1, dtype=torch.int32, device=send.device)
    send_ptr[1:] = torch.cumsum(deg, 0).to(torch.int32)
    return send_ptr, perm.to(torch.int32).contiguous()


def sender_lists(edges, B, N):
    """Cached on the EdgeList: the n_future forwards of one training unroll (train.py:90-112) share their relations."""
    key = (edges.row_ptr.data_ptr(), edges.send.data_ptr(), edges.row_ptr._version, edges.send._version)
    cache = getattr(edges, "_sender_cache", None)
    if cache is None or cache[0] != key:
        with torch.no_grad():
            cache = (key, _sender_lists(edges.row_ptr, edges.send, edges.recv, B, N))
        edges._sender_cache = cache
    return cache[1]


class _DynamicsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, state, attrs, action, p_inst, physics, row_ptr, send, recv, send_ptr, send_perm, *params):
        packed = model.packed_weights()
        F, K = model.nf_effect, model.model_config["pstep"]
        pos, motion, saved = ops.forward_train(packed, state, attrs, action, p_inst, physics, row_ptr, send, recv, F, K)
        ctx.model, ctx.F, ctx.K = model, F, K
        ctx.save_for_backward(packed, state, attrs, action, p_inst, physics, row_ptr, send, recv, saved, motion, send_ptr, send_perm, *params)
        return pos, motion

    @staticmethod
    def backward(ctx, d_pos, d_motion):
        packed, state, attrs, action, p_inst, physics, row_ptr, send, recv, saved, motion, send_ptr, send_perm = ctx.saved_tensors[:13]
        params = ctx.saved_tensors[13:]
        grad_w = [torch.zeros_like(p, dtype=torch.float32).contiguous() for p in params[0::2]]
        grad_b = [torch.zeros_like(p, dtype=torch.float32).contiguous() for p in params[1::2]]
        d_state = torch.zeros_like(state) if ctx.needs_input_grad[1] else None
        d_pos = d_pos.contiguous().float() if d_pos is not None else None
        d_motion = d_motion.contiguous().float() if d_motion is not None else None
        ops.backward(packed, state, attrs, action, p_inst, physics, row_ptr, send, recv, send_ptr, send_perm, saved, motion, d_pos,
                     d_motion, grad_w, grad_b, d_state, ctx.F, ctx.K)
        grads = []
        for w, b in zip(grad_w, grad_b):
            grads += [w, b]
        return (None, d_state, None, None, None, None, None, None, None, None, None, *grads)


def forward_with_grad(model, state, attrs, action, p_instance, physics_param, edges):
    B = state.shape[0]
    n_p = p_instance.shape[1]
    f32 = lambda t: t.to(torch.float32).contiguous()  # noqa: E731
    p_inst = f32(p_instance.reshape(B, n_p, -1)[:, :, 0])
    params = []
    for layer in model._linear_layers():
        params += [layer.weight, layer.bias]
    send_ptr, send_perm = sender_lists(edges, B, state.shape[2])
    return _DynamicsFn.apply(model, f32(state), f32(attrs), f32(action), p_inst, f32(physics_param.to(state.device)), edges.row_ptr,
                             edges.send, edges.recv, send_ptr, send_perm, *params)
