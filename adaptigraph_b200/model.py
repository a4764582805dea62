"""Drop-in `DynamicsPredictor` backed by the sm_100a engine.

Reference: src/dynamics/gnn/model.py:63-313.  Same constructor, same
`forward(state, attrs, Rr, Rs, p_instance, action=None, particle_den=None, obj_mask=None, **kwargs)
-> (pred_pos, pred_motion)`, same 22 state_dict keys (so reference checkpoints load unchanged) and the
same parameter initialisation order (so `torch.manual_seed(s)` + construction yields the reference's
weights).  The sub-modules below only HOLD parameters; all arithmetic runs in the CUDA library through
`ops.forward` — there is no eager fallback.

Additions the reference does not have (optional for callers):
  * `forward(..., edges=EdgeList)` skips the dense one-hot relations entirely;
  * `rollout(...)` runs the autoregressive loop of planning/forward_dynamics.py:156-197 on device.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

import os

from . import _lib as L
from . import ops
from .graph import EdgeList, edges_from_onehots, _thr2_batch


# "fp32": exact FFMA tiles; "tc3": tcgen05 split-fp16, 3 MMAs per product everywhere (fp32-accurate); "tc" (default): the same with
# the relation chain at 2 MMAs per product and the per-relation term stored as 16-bit block fixed point (include/adaptigraph_b200.h)
_PRECISIONS = {"fp32": L.AGX_PREC_FP32, "tc3": L.AGX_PREC_TC_F16X3, "tc": L.AGX_PREC_TC_MIXED}


class _EncoderParams(nn.Module):
    """Parameter holder with the key names of the reference Encoder (model.py:4-21): model.{0,2,4}."""

    def __init__(self, d_in, d_hidden, d_out):
        super().__init__()
        self.model = nn.Sequential(nn.Linear(d_in, d_hidden), nn.ReLU(), nn.Linear(d_hidden, d_hidden), nn.ReLU(),
                                   nn.Linear(d_hidden, d_out), nn.ReLU())

    def layers(self):
        return [self.model[0], self.model[2], self.model[4]]


class _PropagatorParams(nn.Module):
    """Key names of the reference Propagator (model.py:23-41): linear."""

    def __init__(self, d_in, d_out):
        super().__init__()
        self.linear = nn.Linear(d_in, d_out)


class _PredictorParams(nn.Module):
    """Key names of the reference ParticlePredictor (model.py:43-60): linear_{0,1,2}."""

    def __init__(self, d_in, d_hidden, d_out):
        super().__init__()
        self.linear_0 = nn.Linear(d_in, d_hidden)
        self.linear_1 = nn.Linear(d_hidden, d_hidden)
        self.linear_2 = nn.Linear(d_hidden, d_out)


class DynamicsPredictor(nn.Module):
    def __init__(self, model_config, material_config, dataset_config, device):
        super().__init__()
        self.model_config = model_config
        self.material_config = material_config
        self.dataset_config = dataset_config
        self.device = device
        self.n_his = dataset_config["n_his"]
        self.nf_particle = model_config["nf_particle"]
        self.nf_relation = model_config["nf_relation"]
        self.nf_effect = model_config["nf_effect"]
        self.nf_physics = model_config["nf_physics"]
        self.eps = 1e-6
        self.motion_clamp = 100
        # arithmetic of the dense layers (_PRECISIONS above)
        self.precision = _PRECISIONS[os.environ.get("AGX_PRECISION", "tc")]

        self.num_materials = len(material_config["material_index"])
        assert self.num_materials == 1, "Only support single material."
        material_params = material_config[dataset_config["materials"][0]]["physics_params"]
        self.material_dim = sum(1 for p in material_params if p["use"])

        mc = model_config
        unsupported = {k: mc[k] for k in ("state_dim", "offset_dim", "density_dim", "rel_density_dim") if mc[k] != 0}
        if mc["rel_particle_dim"] not in (0,):
            unsupported["rel_particle_dim"] = mc["rel_particle_dim"]
        if mc["rel_attr_dim"] != mc["attr_dim"] or mc["rel_group_dim"] != 1 or mc["rel_distance_dim"] != 3:
            unsupported["rel_*"] = (mc["rel_attr_dim"], mc["rel_group_dim"], mc["rel_distance_dim"])
        if not (self.nf_particle == self.nf_relation == self.nf_effect):
            unsupported["nf_*"] = (self.nf_particle, self.nf_relation, self.nf_effect)
        if unsupported:
            raise NotImplementedError(
                "adaptigraph_b200 implements the feature layout of the shipped configs "
                "(config/dynamics/{rope,granular,cloth}.yaml:55-78); unsupported settings: " + repr(unsupported))

        input_dim = mc["attr_dim"] + mc["action_dim"] + self.material_dim                     # model.py:96-101
        rel_input_dim = mc["rel_attr_dim"] * 2 + mc["rel_group_dim"] + mc["rel_distance_dim"] * self.n_his  # :109-113
        # construction order = the reference's (model.py:103-122), so seeded init matches
        self.particle_encoder = _EncoderParams(input_dim, self.nf_particle, self.nf_effect)
        self.relation_encoder = _EncoderParams(rel_input_dim, self.nf_relation, self.nf_effect)
        self.particle_propagator = _PropagatorParams(self.nf_effect * 2, self.nf_effect)
        self.relation_propagator = _PropagatorParams(self.nf_effect * 3, self.nf_effect)
        self.non_rigid_predictor = _PredictorParams(self.nf_effect, self.nf_effect, 3)
        self._packed = None
        self._packed_key = None
        if mc["verbose"]:
            print("DynamicsPredictor initialized")
            print("particle input dim: {}, relation input dim: {}".format(input_dim, rel_input_dim))

    def set_precision(self, name: str) -> "DynamicsPredictor":
        """'fp32' (exact FFMA tiles), 'tc3' (tcgen05 tensor cores on split-fp16 operands, fp32-accurate) or 'tc' (default: tc3 with
        the relation chain's precision budget spent, rollout RMSE ~3.5e-6 against the 1e-4 tolerance)."""
        self.precision = _PRECISIONS[name]
        return self

    # ------------------------------------------------------------------ weights
    def _linear_layers(self):
        return (self.particle_encoder.layers() + self.relation_encoder.layers() +
                [self.particle_propagator.linear, self.relation_propagator.linear,
                 self.non_rigid_predictor.linear_0, self.non_rigid_predictor.linear_1, self.non_rigid_predictor.linear_2])

    def packed_weights(self) -> torch.Tensor:
        """Packed device blob for the kernels, rebuilt when any parameter was modified or moved."""
        layers = self._linear_layers()
        key = tuple((p.data_ptr(), p._version) for l in layers for p in (l.weight, l.bias))
        if self._packed is None or key != self._packed_key:
            mc = self.model_config
            with torch.no_grad():
                self._packed = ops.pack_weights([l.weight.detach() for l in layers], [l.bias.detach() for l in layers],
                                                self.n_his, mc["attr_dim"], self.material_dim, mc["action_dim"])
            self._packed_key = key
        return self._packed

    # ------------------------------------------------------------------ forward
    def forward(self, state, attrs, Rr=None, Rs=None, p_instance=None, action=None, particle_den=None, obj_mask=None,
                edges: Optional[EdgeList] = None, **kwargs):
        physics_keys = [k for k in kwargs.keys() if k.endswith("_physics_param")]
        assert len(physics_keys) == 1                                                       # model.py:184-185
        physics_param = kwargs[physics_keys[0]]
        assert action is not None                                                           # model.py:193-194
        assert p_instance is not None
        if p_instance.dim() == 3 and p_instance.shape[2] != 1:
            raise NotImplementedError("p_instance with more than one instance column (max_n > 1) is not supported")
        if torch.is_grad_enabled() and (state.requires_grad or any(p.requires_grad for p in self.parameters())) \
                and not getattr(self, "_allow_no_backward", False):
            from .autograd import forward_with_grad  # noqa: WPS433  (raises if the backward kernels are unavailable)
            if edges is None:
                edges = edges_from_onehots(Rr, Rs)
            return forward_with_grad(self, state, attrs, action, p_instance, physics_param, edges)
        if edges is None:
            if Rr is None or Rs is None:
                raise ValueError("forward needs either dense Rr/Rs or edges=EdgeList")
            edges = edges_from_onehots(Rr, Rs)
        physics_param = physics_param.to(state.device)                                      # reference builds it on CPU in planning
        return ops.forward(self.packed_weights(), state, attrs, action, p_instance, physics_param,
                           edges.row_ptr, edges.send, edges.recv, self.nf_effect, self.model_config["pstep"],
                           self.precision)

    # ------------------------------------------------------------------ rollout
    @torch.no_grad()
    def rollout(self, state, attrs, action, p_instance, physics_param, state_mask, eef_mask, adj_thresh, topk,
                connect_tools_all, n_steps, max_nR, y_mode="min", gripper_raise=0.0, check=True):
        """Device-resident autoregressive rollout with per-step re-graphing
        (planning/forward_dynamics.py:156-197; y_mode='masked_mean' gives :351-393).

        state (B,H,N,3) is NOT modified (a copy is advanced).  Returns a dict with
        'state_seqs' (B, n_steps, n_p, 3), 'n_edges' (n_steps, B) and the final history 'state'.
        max_nR is the per-graph relation capacity of the reference configs; exceeding it raises
        like pad_torch (utils.py:37-46) when check=True (one host sync at the end).
        """
        B, H, N, _ = state.shape
        hist = state.detach().to(torch.float32).contiguous().clone()
        thr2 = _thr2_batch(adj_thresh, B, state.device)
        mode = {"min": L.AGX_Y_MIN, "masked_mean": L.AGX_Y_MASKED_MEAN}[y_mode]
        pred_seq, n_edges, status = ops.rollout(
            self.packed_weights(), hist, attrs, action, p_instance, physics_param.to(state.device), state_mask, eef_mask,
            thr2, self.nf_effect, self.model_config["pstep"], topk, connect_tools_all, n_steps, mode,
            float(gripper_raise), B * max_nR, self.precision)
        if check:
            worst = int(n_edges.max().item())
            if int(status.item()) & 1 or worst > max_nR:
                raise RuntimeError(f"rollout: a graph reached {worst} relations, capacity max_nR={max_nR}")
        return {"state_seqs": pred_seq, "n_edges": n_edges, "state": hist, "status": status}


class GraphedRollout:
    """`DynamicsPredictor.rollout` captured once as a CUDA graph and replayed: one launch instead of ~17 per model step, which is
    what an MPC loop with small sample batches is bound by.  Shapes, relation capacity, step count and the scalar arguments are
    frozen at construction; every call copies the new tensors into the captured buffers and replays.

        roll = GraphedRollout(model, state, attrs, action, p_instance, physics_param, state_mask, eef_mask,
                              adj_thresh, topk, connect_tools_all, n_steps, max_nR)
        out = roll(state=new_state, action=new_action)      # any subset of the tensor arguments; returns the captured outputs

    The returned tensors are overwritten by the next call.  Capacity overflow is reported by `out["status"]` / `check()`
    (no host synchronisation inside a replay)."""

    _TENSORS = ("state", "attrs", "action", "p_instance", "physics_param", "state_mask", "eef_mask")

    def __init__(self, model, state, attrs, action, p_instance, physics_param, state_mask, eef_mask, adj_thresh, topk,
                 connect_tools_all, n_steps, max_nR, y_mode="min", gripper_raise=0.0):
        if not state.is_cuda:
            raise RuntimeError("GraphedRollout needs CUDA tensors (no CPU path)")
        self.model, self.max_nR = model, max_nR
        self.static = {k: v.detach().clone() for k, v in zip(self._TENSORS, (state, attrs, action, p_instance,
                                                                            physics_param.to(state.device), state_mask, eef_mask))}
        self.static["state"] = self.static["state"].to(torch.float32).contiguous()
        B, dev = state.shape[0], state.device
        thr2 = _thr2_batch(adj_thresh, B, dev)
        mode = {"min": L.AGX_Y_MIN, "masked_mean": L.AGX_Y_MASKED_MEAN}[y_mode]
        packed = model.packed_weights()

        hist = torch.empty_like(self.static["state"])      # the history the rollout advances in place

        def run():
            hist.copy_(self.static["state"])
            pred_seq, n_edges, status = ops.rollout(
                packed, hist, self.static["attrs"], self.static["action"], self.static["p_instance"], self.static["physics_param"],
                self.static["state_mask"], self.static["eef_mask"], thr2, model.nf_effect, model.model_config["pstep"], topk,
                connect_tools_all, n_steps, mode, float(gripper_raise), B * max_nR, model.precision)
            return {"state_seqs": pred_seq, "n_edges": n_edges, "state": hist, "status": status}

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(2):
                run()
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.out = run()
        self._packed_key = model._packed_key
        self._keep = (thr2, packed, hist)      # the graph holds raw addresses: everything it reads must outlive __init__

    def __call__(self, **tensors):
        # packed_weights() first: it is what refreshes the key (and the blob) after an in-place parameter change
        if self.model.packed_weights() is not self._keep[1] or self.model._packed_key != self._packed_key:
            raise RuntimeError("the model's parameters changed since capture: build a new GraphedRollout")
        for k, v in tensors.items():
            if k not in self.static:
                raise KeyError(f"{k} is not a tensor argument of the captured rollout ({', '.join(self._TENSORS)})")
            if v.shape != self.static[k].shape:
                raise RuntimeError(f"GraphedRollout was captured for {k} of shape {tuple(self.static[k].shape)}, got {tuple(v.shape)}")
            self.static[k].copy_(v, non_blocking=True)      # pinned host tensors: stream-ordered H2D, no host synchronisation
        self.graph.replay()
        return self.out

    def check(self):
        """Synchronises; raises like pad_torch (utils.py:37-46) if a graph exceeded max_nR in the last replay."""
        worst = int(self.out["n_edges"].max().item())
        if int(self.out["status"].item()) & 1 or worst > self.max_nR:
            raise RuntimeError(f"rollout: a graph reached {worst} relations, capacity max_nR={self.max_nR}")
        return self.out
