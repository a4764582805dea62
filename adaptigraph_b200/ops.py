"""PyTorch custom ops over the C ABI (include/adaptigraph_b200.h).

Each op validates shapes/dtypes/devices on the Python side, allocates outputs
and the workspace from torch's caching allocator, and calls the library on the
current CUDA stream through ctypes.  Nothing here computes: a CPU tensor is an
error, not a fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Tuple

import torch
from torch import Tensor

from . import _lib as L

lib = L.lib


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Tensor) -> C.c_void_p:
    return C.c_void_p(t.data_ptr())


def _need_cuda(*ts: Tensor) -> None:
    for t in ts:
        if not t.is_cuda:
            raise RuntimeError("adaptigraph_b200 runs on CUDA tensors only (no CPU fallback); got a "
                               f"{t.device} tensor of shape {tuple(t.shape)}")


def _f32(t: Tensor) -> Tensor:
    return t.to(torch.float32).contiguous()


def _u8(t: Tensor) -> Tensor:
    return t.to(torch.uint8).contiguous()


def make_dims(F: int, n_his: int, d_attr: int, d_phys: int, d_act: int, pstep: int) -> L.AgxModelDims:
    return L.AgxModelDims(F, n_his, d_attr, d_phys, d_act, pstep)


def launch_count() -> int:
    return int(lib.agx_launch_count())


# --------------------------------------------------------------------------- weights
@torch.library.custom_op("agx::pack_weights", mutates_args=())
def pack_weights(weights: List[Tensor], biases: List[Tensor], n_his: int, d_attr: int, d_phys: int, d_act: int) -> Tensor:
    """Pads, transposes and splits the 11 reference-layout Linear layers into the packed blob the kernels read."""
    _need_cuda(*weights, *biases)
    assert len(weights) == L.AGX_NUM_LAYERS and len(biases) == L.AGX_NUM_LAYERS
    F = weights[1].shape[0]
    dims = make_dims(F, n_his, d_attr, d_phys, d_act, 1)
    ws = [_f32(w) for w in weights]
    bs = [_f32(b) for b in biases]
    raw = L.AgxWeights()
    for i in range(L.AGX_NUM_LAYERS):
        raw.weight[i] = ws[i].data_ptr()
        raw.bias[i] = bs[i].data_ptr()
    n = lib.agx_packed_weights_bytes(C.byref(dims))
    out = torch.empty(n // 4, dtype=torch.float32, device=ws[0].device)
    L.check(lib.agx_pack_weights(C.byref(dims), C.byref(raw), _ptr(out), _stream()), "agx_pack_weights")
    return out


@pack_weights.register_fake
def _(weights, biases, n_his, d_attr, d_phys, d_act):
    return weights[0].new_empty(1)


# --------------------------------------------------------------------------- graph construction
@torch.library.custom_op("agx::graph_build", mutates_args=())
def graph_build(pos: Tensor, mask: Tensor, tool_mask: Tensor, thr2: Tensor, topk: int, connect_tools_all: bool,
                semantics: int, cap: int) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    """(B,N,3) positions -> CSR relations: row_ptr (B*N+1), send (cap), recv (cap), n_edges (B), status (1)."""
    _need_cuda(pos, mask, tool_mask, thr2)
    B, N, _ = pos.shape
    pos, thr2 = _f32(pos), _f32(thr2)
    mask, tool_mask = _u8(mask), _u8(tool_mask)
    dev = pos.device
    row_ptr = torch.empty(B * N + 1, dtype=torch.int32, device=dev)
    send = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)
    recv = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)
    n_edges = torch.empty(B, dtype=torch.int32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    nws = lib.agx_graph_workspace_bytes(B, N, topk)
    ws = torch.empty(nws, dtype=torch.uint8, device=dev)
    L.check(lib.agx_graph_build(_ptr(pos), _ptr(mask), _ptr(tool_mask), _ptr(thr2), B, N, topk, int(connect_tools_all),
                                semantics, _ptr(row_ptr), _ptr(send), _ptr(recv), cap, _ptr(n_edges), _ptr(status),
                                _ptr(ws), nws, _stream()), "agx_graph_build")
    return row_ptr, send, recv, n_edges, status


@graph_build.register_fake
def _(pos, mask, tool_mask, thr2, topk, connect_tools_all, semantics, cap):
    B, N, _ = pos.shape
    i32 = dict(dtype=torch.int32, device=pos.device)
    return (torch.empty(B * N + 1, **i32), torch.empty(max(cap, 1), **i32), torch.empty(max(cap, 1), **i32),
            torch.empty(B, **i32), torch.empty(1, **i32))


@torch.library.custom_op("agx::onehot_to_ids", mutates_args=())
def onehot_to_ids(R: Tensor) -> Tensor:
    """Dense one-hot relation rows (B, n_rel, N) -> int32 (B, n_rel) particle id per row, -1 for all-zero rows."""
    _need_cuda(R)
    R = _f32(R)
    B, n_rel, N = R.shape
    ids = torch.empty(B, n_rel, dtype=torch.int32, device=R.device)
    if n_rel:
        L.check(lib.agx_onehot_to_ids(_ptr(R), B, n_rel, N, _ptr(ids), _stream()), "agx_onehot_to_ids")
    return ids


@onehot_to_ids.register_fake
def _(R):
    return R.new_empty(R.shape[:2], dtype=torch.int32)


@torch.library.custom_op("agx::edges_to_onehot", mutates_args=())
def edges_to_onehot(row_ptr: Tensor, send: Tensor, B: int, N: int, n_rel: int) -> Tuple[Tensor, Tensor]:
    """CSR relations -> the reference's dense (B, n_rel, N) Rr, Rs (graph.py:152-155)."""
    _need_cuda(row_ptr, send)
    Rr = torch.zeros(B, n_rel, N, dtype=torch.float32, device=row_ptr.device)
    Rs = torch.zeros(B, n_rel, N, dtype=torch.float32, device=row_ptr.device)
    if n_rel:
        L.check(lib.agx_edges_to_onehot(_ptr(row_ptr), _ptr(send), B, N, n_rel, _ptr(Rr), _ptr(Rs), _stream()),
                "agx_edges_to_onehot")
    return Rr, Rs


@edges_to_onehot.register_fake
def _(row_ptr, send, B, N, n_rel):
    return (row_ptr.new_empty((B, n_rel, N), dtype=torch.float32), row_ptr.new_empty((B, n_rel, N), dtype=torch.float32))


# --------------------------------------------------------------------------- particle sampling
@torch.library.custom_op("agx::fps", mutates_args=())
def fps(pos: Tensor, n_points: Tensor, start_idx: Tensor, max_samples: int, radius: float) -> Tuple[Tensor, Tensor]:
    """Farthest-point sampling of B clouds: pos (B,N,3), n_points (B) int32 valid prefix lengths, start_idx (B) int32.
    radius < 0: exactly max_samples picks (dgl.geometry.farthest_point_sampler); radius >= 0: picks until every point lies
    within `radius` of a pick (fps_rad_idx, utils.py:10-24).  Returns idx (B, max_samples) int32 and counts (B) int32."""
    _need_cuda(pos, n_points, start_idx)
    pos = _f32(pos)
    B, N, _ = pos.shape
    n_points = n_points.to(torch.int32).contiguous()
    start_idx = start_idx.to(torch.int32).contiguous()
    idx = torch.zeros(B, max_samples, dtype=torch.int32, device=pos.device)
    cnt = torch.empty(B, dtype=torch.int32, device=pos.device)
    L.check(lib.agx_fps(_ptr(pos), _ptr(n_points), B, N, max_samples, _ptr(start_idx), float(radius), _ptr(idx), _ptr(cnt),
                        _stream()), "agx_fps")
    return idx, cnt


@fps.register_fake
def _(pos, n_points, start_idx, max_samples, radius):
    B = pos.shape[0]
    return pos.new_empty((B, max_samples), dtype=torch.int32), pos.new_empty((B,), dtype=torch.int32)


def fps_radii(pos: Tensor, n_points: Tensor, start_idx: Tensor, max_samples: int, radii: Tensor) -> Tuple[Tensor, Tensor]:
    """`fps` in its radius form with one radius per cloud: radii (B) float64 (agx_fps_radii)."""
    _need_cuda(pos, n_points, start_idx, radii)
    pos = _f32(pos)
    B, N, _ = pos.shape
    n_points = n_points.to(torch.int32).contiguous()
    start_idx = start_idx.to(torch.int32).contiguous()
    radii = radii.to(torch.float64).contiguous()
    if radii.numel() != B or bool((radii < 0).any()):
        raise ValueError("fps_radii: one non-negative radius per cloud")
    idx = torch.zeros(B, max_samples, dtype=torch.int32, device=pos.device)
    cnt = torch.empty(B, dtype=torch.int32, device=pos.device)
    L.check(lib.agx_fps_radii(_ptr(pos), _ptr(n_points), B, N, max_samples, _ptr(start_idx), _ptr(radii), _ptr(idx), _ptr(cnt),
                              _stream()), "agx_fps_radii")
    return idx, cnt


@torch.library.custom_op("agx::chamfer", mutates_args=())
def chamfer(x: Tensor, y: Tensor) -> Tensor:
    """planning/losses.py:4-10: x (B,N,3), y (1,M,3) or (B,M,3) -> (B) symmetric mean-of-min distances."""
    _need_cuda(x, y)
    x, y = _f32(x), _f32(y)
    B, N, D = x.shape
    By, M, Dy = y.shape
    if D != 3 or Dy != 3 or By not in (1, B):
        raise ValueError(f"chamfer: expected x (B,N,3) and y (1|B,M,3), got {tuple(x.shape)} and {tuple(y.shape)}")
    out = torch.empty(B, dtype=torch.float32, device=x.device)
    L.check(lib.agx_chamfer(_ptr(x), _ptr(y), B, N, M, int(By == B and B > 1), _ptr(out), _stream()), "agx_chamfer")
    return out


@chamfer.register_fake
def _(x, y):
    return x.new_empty((x.shape[0],))


_RC_WS = {}     # (device index, bytes) -> zero-initialised workspace of agx_running_cost (the kernel re-zeroes its counters itself)


@torch.library.custom_op("agx::running_cost", mutates_args=())
def running_cost(state: Tensor, action: Tensor, state_cur: Tensor, bbox: Tensor, target: Tensor, error_mode: int,
                 penalty_mode: int, sim_real_ratio: float) -> Tensor:
    """planning/plan.py:27-59 with its error / penalty terms (losses.py) in one launch: state (bsz,L,n,3), action (bsz,L,A>=3),
    state_cur (n,3), bbox (2,2), target (M,3) points [error_mode 0: chamfer] or (2,2) box [1: box_loss] -> reward_seqs (bsz)."""
    _need_cuda(state, action, state_cur, bbox, target)
    state, action, state_cur, bbox, target = _f32(state), _f32(action), _f32(state_cur), _f32(bbox), _f32(target)
    bsz, Lk, n, D = state.shape
    if D != 3 or action.shape[:2] != (bsz, Lk) or state_cur.shape != (n, 3) or bbox.shape != (2, 2):
        raise ValueError(f"running_cost: bad shapes state {tuple(state.shape)} action {tuple(action.shape)} state_cur "
                         f"{tuple(state_cur.shape)} bbox {tuple(bbox.shape)}")
    if error_mode == L.AGX_ERROR_CHAMFER:
        if target.dim() != 2 or target.shape[1] != 3:
            raise ValueError(f"running_cost: chamfer target must be (M,3), got {tuple(target.shape)}")
        M = target.shape[0]
    else:
        if target.shape != (2, 2):
            raise ValueError(f"running_cost: box target must be (2,2), got {tuple(target.shape)}")
        M = 0
    nws = int(lib.agx_running_cost_workspace_bytes(bsz, Lk))
    key = (state.device.index, nws)
    ws = _RC_WS.get(key)
    if ws is None:
        ws = _RC_WS[key] = torch.zeros(nws, dtype=torch.uint8, device=state.device)
    out = torch.empty(bsz, dtype=torch.float32, device=state.device)
    L.check(lib.agx_running_cost(_ptr(state), _ptr(action), action.shape[2], _ptr(state_cur), _ptr(bbox), error_mode, _ptr(target), M,
                                 penalty_mode, float(sim_real_ratio), bsz, Lk, n, _ptr(ws), nws, _ptr(out), _stream()), "agx_running_cost")
    return out


@running_cost.register_fake
def _(state, action, state_cur, bbox, target, error_mode, penalty_mode, sim_real_ratio):
    return state.new_empty((state.shape[0],))


# --------------------------------------------------------------------------- forward / rollout
@torch.library.custom_op("agx::forward", mutates_args=())
def forward(packed: Tensor, state: Tensor, attrs: Tensor, action: Tensor, p_instance: Tensor, physics: Tensor,
            row_ptr: Tensor, send: Tensor, recv: Tensor, F: int, pstep: int, precision: int) -> Tuple[Tensor, Tensor]:
    """DynamicsPredictor.forward on CSR relations -> (pred_pos, pred_motion), each (B, n_p, 3)."""
    _need_cuda(packed, state, attrs, action, p_instance, physics, row_ptr, send, recv)
    B, H, N, _ = state.shape
    n_p = p_instance.shape[1]
    state, attrs, action, physics = _f32(state), _f32(attrs), _f32(action), _f32(physics)
    p_inst = _f32(p_instance.reshape(B, n_p, -1)[:, :, 0])
    dims = make_dims(F, H, attrs.shape[2], physics.shape[1], action.shape[2], pstep)
    E_cap = int(send.numel())
    g = L.AgxGraphIn(B, N, n_p, state.data_ptr(), attrs.data_ptr(), action.data_ptr(), p_inst.data_ptr(),
                     physics.data_ptr(), row_ptr.data_ptr(), send.data_ptr(), recv.data_ptr(), E_cap)
    dev = state.device
    pred_pos = torch.empty(B, n_p, 3, dtype=torch.float32, device=dev)
    pred_motion = torch.empty(B, n_p, 3, dtype=torch.float32, device=dev)
    nws = lib.agx_forward_workspace_bytes(C.byref(dims), B, N, E_cap)
    ws = torch.empty(nws, dtype=torch.uint8, device=dev)
    L.check(lib.agx_forward(C.byref(dims), _ptr(packed), C.byref(g), _ptr(pred_pos), n_p * 3, _ptr(pred_motion),
                            precision, _ptr(ws), nws, _stream()), "agx_forward")
    return pred_pos, pred_motion


@forward.register_fake
def _(packed, state, attrs, action, p_instance, physics, row_ptr, send, recv, F, pstep, precision):
    B, n_p = p_instance.shape[0], p_instance.shape[1]
    return state.new_empty((B, n_p, 3)), state.new_empty((B, n_p, 3))


@torch.library.custom_op("agx::rollout", mutates_args=("state",))
def rollout(packed: Tensor, state: Tensor, attrs: Tensor, action: Tensor, p_instance: Tensor, physics: Tensor,
            mask: Tensor, tool_mask: Tensor, thr2: Tensor, F: int, pstep: int, topk: int, connect_tools_all: bool,
            n_steps: int, y_mode: int, gripper_raise: float, E_cap: int, precision: int) -> Tuple[Tensor, Tensor, Tensor]:
    """T autoregressive steps on device; `state` (B,H,N,3) is advanced in place.
    Returns pred_seq (B,T,n_p,3), n_edges (T,B) int32, status (1) int32."""
    _need_cuda(packed, state, attrs, action, p_instance, physics, mask, tool_mask, thr2)
    if state.dtype != torch.float32 or not state.is_contiguous():
        raise ValueError("rollout: state must be a contiguous float32 tensor (it is updated in place)")
    B, H, N, _ = state.shape
    n_p = p_instance.shape[1]
    attrs, action, physics, thr2 = _f32(attrs), _f32(action), _f32(physics), _f32(thr2)
    p_inst = _f32(p_instance.reshape(B, n_p, -1)[:, :, 0])
    mask, tool_mask = _u8(mask), _u8(tool_mask)
    dims = make_dims(F, H, attrs.shape[2], physics.shape[1], action.shape[2], pstep)
    r = L.AgxRolloutIn(B, N, n_p, state.data_ptr(), attrs.data_ptr(), action.data_ptr(), p_inst.data_ptr(),
                       physics.data_ptr(), mask.data_ptr(), tool_mask.data_ptr(), thr2.data_ptr(), topk,
                       int(connect_tools_all), n_steps, y_mode, gripper_raise, E_cap)
    dev = state.device
    pred_seq = torch.empty(B, n_steps, n_p, 3, dtype=torch.float32, device=dev)
    n_edges = torch.empty(n_steps, B, dtype=torch.int32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    nws = lib.agx_rollout_workspace_bytes(C.byref(dims), B, N, E_cap, topk)
    ws = torch.empty(nws, dtype=torch.uint8, device=dev)
    L.check(lib.agx_rollout(C.byref(dims), _ptr(packed), C.byref(r), _ptr(pred_seq), _ptr(n_edges), _ptr(status),
                            precision, _ptr(ws), nws, _stream()), "agx_rollout")
    return pred_seq, n_edges, status


@rollout.register_fake
def _(packed, state, attrs, action, p_instance, physics, mask, tool_mask, thr2, F, pstep, topk, connect_tools_all,
      n_steps, y_mode, gripper_raise, E_cap, precision):
    B, n_p = p_instance.shape[0], p_instance.shape[1]
    return (state.new_empty((B, n_steps, n_p, 3)), state.new_empty((n_steps, B), dtype=torch.int32),
            state.new_empty((1,), dtype=torch.int32))


# --------------------------------------------------------------------------- per-kernel timing (bench.py)
def profile_enable(on: bool) -> None:
    L.check(lib.agx_profile_enable(int(on)), "agx_profile_enable")


def profile_read() -> dict:
    """{kernel kind: (total ms, launches)} accumulated since the last read (synchronises on the recorded events)."""
    ms = (C.c_double * L.AGX_NUM_KINDS)()
    cnt = (C.c_int64 * L.AGX_NUM_KINDS)()
    L.check(lib.agx_profile_read(ms, cnt), "agx_profile_read")
    return {lib.agx_kind_name(i).decode(): (ms[i], cnt[i]) for i in range(L.AGX_NUM_KINDS) if cnt[i]}


# --------------------------------------------------------------------------- training (forward with saved activations, backward)
def _graph_in(state, attrs, action, p_inst, physics, row_ptr, send, recv):
    B, H, N, _ = state.shape
    n_p = p_inst.shape[1]
    return L.AgxGraphIn(B, N, n_p, state.data_ptr(), attrs.data_ptr(), action.data_ptr(), p_inst.data_ptr(), physics.data_ptr(),
                        row_ptr.data_ptr(), send.data_ptr(), recv.data_ptr(), int(send.numel()))


def forward_train(packed: Tensor, state: Tensor, attrs: Tensor, action: Tensor, p_inst: Tensor, physics: Tensor, row_ptr: Tensor,
                  send: Tensor, recv: Tensor, F: int, pstep: int):
    """Exact-fp32 forward that keeps every activation the backward needs.  Inputs must already be contiguous fp32 / int32
    CUDA tensors (p_inst is (B, n_p)).  Returns pred_pos, pred_motion, saved (opaque uint8 buffer)."""
    _need_cuda(packed, state, attrs, action, p_inst, physics, row_ptr, send, recv)
    B, H, N, _ = state.shape
    n_p = p_inst.shape[1]
    dims = make_dims(F, H, attrs.shape[2], physics.shape[1], action.shape[2], pstep)
    g = _graph_in(state, attrs, action, p_inst, physics, row_ptr, send, recv)
    dev = state.device
    pred_pos = torch.empty(B, n_p, 3, dtype=torch.float32, device=dev)
    pred_motion = torch.empty(B, n_p, 3, dtype=torch.float32, device=dev)
    nsv = lib.agx_train_saved_bytes(C.byref(dims), B, N, int(send.numel()))
    saved = torch.empty(nsv, dtype=torch.uint8, device=dev)
    L.check(lib.agx_forward_train(C.byref(dims), _ptr(packed), C.byref(g), _ptr(pred_pos), n_p * 3, _ptr(pred_motion), _ptr(saved), nsv,
                                  _stream()), "agx_forward_train")
    return pred_pos, pred_motion, saved


def train_saved_views(saved: Tensor, F: int, n_his: int, d_attr: int, d_phys: int, d_act: int, pstep: int, B: int, N: int, E_cap: int):
    """Row-major views / copies ([rows][AGX_FP] fp32) of the forward activations inside the opaque `saved` buffer of
    `forward_train`, keyed h1, h2, penc, g1, g2, renc, C, u1, u2 and P<k+1>, agg<k>, Qr<k>, Qs<k> per propagation step k.
    For parity tests (ReLU activity patterns); the blocked C / Qr / Qs of the tensor-core forward are un-blocked here."""
    dims = make_dims(F, n_his, d_attr, d_phys, d_act, pstep)
    n = 9 + 4 * pstep
    offs = (C.c_int64 * n)()
    blocked = lib.agx_train_saved_offsets(C.byref(dims), B, N, E_cap, offs, n)
    if blocked < 0:
        L.check(blocked, "agx_train_saved_offsets")
    rows, E = B * N, max(E_cap, 1)
    fl = saved.view(torch.float32)

    def row_major(off, r):
        return fl[off // 4: off // 4 + r * L.AGX_FP].view(r, L.AGX_FP)

    def unblock(off, r):
        tiles = (r + 127) // 128
        t = fl[off // 4: off // 4 + tiles * 128 * 152].view(tiles, 128 * 152)
        wide = t[:, : 9 * 128 * 16].reshape(tiles, 9, 128, 16).permute(0, 2, 1, 3).reshape(tiles, 128, 144)
        narrow = t[:, 9 * 128 * 16:].reshape(tiles, 128, 8)
        out = torch.zeros(tiles * 128, L.AGX_FP, dtype=torch.float32, device=saved.device)
        out[:, :152] = torch.cat([wide, narrow], 2).reshape(tiles * 128, 152)
        return out[:r]

    feat = unblock if blocked == 1 else row_major
    names = ["h1", "h2", "penc", "g1", "g2", "renc", "C", "u1", "u2"]
    nrow = [rows, rows, rows, E, E, E, E, rows, rows]
    views = {}
    for i, (k, r) in enumerate(zip(names, nrow)):
        views[k] = feat(offs[i], r) if k == "C" else row_major(offs[i], r)
    for k in range(pstep):
        views[f"P{k + 1}"] = row_major(offs[9 + 4 * k], rows)
        views[f"agg{k}"] = row_major(offs[10 + 4 * k], rows)
        views[f"Qr{k}"] = feat(offs[11 + 4 * k], rows)
        views[f"Qs{k}"] = feat(offs[12 + 4 * k], rows)
    return views


def backward(packed: Tensor, state: Tensor, attrs: Tensor, action: Tensor, p_inst: Tensor, physics: Tensor, row_ptr: Tensor,
             send: Tensor, recv: Tensor, send_ptr: Tensor, send_perm: Tensor, saved: Tensor, pred_motion: Tensor, d_pos, d_motion,
             grad_w: List[Tensor], grad_b: List[Tensor], d_state, F: int, pstep: int) -> None:
    """Accumulates parameter gradients into grad_w / grad_b (reference layout, fp32, contiguous) and into d_state (nullable)."""
    B, H, N, _ = state.shape
    dims = make_dims(F, H, attrs.shape[2], physics.shape[1], action.shape[2], pstep)
    g = _graph_in(state, attrs, action, p_inst, physics, row_ptr, send, recv)
    wg = L.AgxWeightGrads()
    for i in range(L.AGX_NUM_LAYERS):
        wg.weight[i] = grad_w[i].data_ptr()
        wg.bias[i] = grad_b[i].data_ptr()
    nsc = lib.agx_train_scratch_bytes(C.byref(dims), B, N, int(send.numel()))
    scratch = torch.empty(nsc, dtype=torch.uint8, device=state.device)
    nul = C.c_void_p(None)
    L.check(lib.agx_backward(C.byref(dims), _ptr(packed), C.byref(g), _ptr(saved), _ptr(send_ptr), _ptr(send_perm), _ptr(pred_motion),
                             _ptr(d_pos) if d_pos is not None else nul, _ptr(d_motion) if d_motion is not None else nul, C.byref(wg),
                             _ptr(d_state) if d_state is not None else nul, _ptr(scratch), nsc, _stream()), "agx_backward")


# --------------------------------------------------------------------------- optimiser
def adam_step(params: Tensor, grads: Tensor, exp_avg: Tensor, exp_avg_sq: Tensor, step: Tensor, lr: float, beta1: float, beta2: float,
              eps: float, grad_scale: float = 1.0) -> None:
    """In-place Adam over flat fp32 buckets (train.py:63 semantics); `step` is a device int32 counter, incremented on the stream."""
    _need_cuda(params, grads, exp_avg, exp_avg_sq, step)
    for t in (params, grads, exp_avg, exp_avg_sq):
        if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != params.numel():
            raise ValueError("adam_step: buckets must be contiguous float32 tensors of equal length")
    if step.dtype != torch.int32 or step.numel() != 1:
        raise ValueError("adam_step: step must be a one-element int32 tensor")
    L.check(lib.agx_adam_step(_ptr(params), _ptr(grads), _ptr(exp_avg), _ptr(exp_avg_sq), params.numel(), lr, beta1, beta2, eps,
                              grad_scale, _ptr(step), _stream()), "agx_adam_step")
