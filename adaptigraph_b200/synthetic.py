"""Seeded synthetic rope / granular / cloth particle graphs.

These are the inputs every parity test, golden fixture and bench run uses.  The
reference ships no data (checkpoints and episodes live behind Google-Drive
links), so the workloads are generated: particle layouts that reproduce the
edge densities of the reference's own configs

* rope     : adj_thresh 0.5,  topk 10, connect_tools_all False  (config/planning/rope.yaml:10-14)
* granular : adj_thresh 0.4,  topk 20, connect_tools_all False  (config/planning/granular.yaml:10-14)
* cloth    : adj_thresh 0.75, topk 5,  connect_tools_all True   (config/dynamics/cloth.yaml:28-31)

All positions carry N(0, 0.01^2) noise so that no two pair distances tie
(torch.topk tie order is unspecified; the reference inherits that).

Everything here is CPU torch; callers move tensors to the device.
"""
from __future__ import annotations

import copy
import math
from dataclasses import dataclass
from typing import Dict

import torch

MATERIALS = {
    # name: (adj_thresh, topk, connect_tools_all, n_tool)
    "rope": (0.5, 10, False, 1),
    "granular": (0.4, 20, False, 5),
    "cloth": (0.75, 5, True, 2),
}


def model_config(pstep: int = 3, nf: int = 150) -> dict:
    """`model_config` section shared by config/dynamics/{rope,granular,cloth}.yaml:55-78."""
    return {
        "verbose": False,
        "nf_particle": nf, "nf_relation": nf, "nf_effect": nf, "nf_physics": 10,
        "attr_dim": 2, "state_dim": 0, "offset_dim": 0, "action_dim": 3, "density_dim": 0,
        "pstep": pstep, "sequence_len": 4,
        "rel_particle_dim": 0, "rel_attr_dim": 2, "rel_group_dim": 1,
        "rel_distance_dim": 3, "rel_density_dim": 0,
    }


def material_config(material: str) -> dict:
    """One physics parameter in use, like every shipped material_config (rope.yaml:84-112)."""
    return {
        "material_index": {material: 0},
        material: {"physics_params": [
            {"name": "unused", "use": False, "min": 0.0, "max": 1.0},
            {"name": "stiffness", "use": True, "min": 0.0, "max": 1.0},
        ]},
    }


def dataset_config(material: str, n_his: int = 4) -> dict:
    return {"data_name": material, "materials": [material], "n_his": n_his, "n_future": 3}


def configs(material: str, pstep: int = 3, nf: int = 150, n_his: int = 4):
    return (model_config(pstep, nf), material_config(material), dataset_config(material, n_his))


@dataclass
class Workload:
    """A batch of particle graphs, tensors named as the reference's graph dict
    (planning/forward_dynamics.py:130-147)."""
    material: str
    state: torch.Tensor        # (B, H, N, 3) f32
    attrs: torch.Tensor        # (B, N, 2) f32
    action: torch.Tensor       # (B, N, 3) f32
    p_instance: torch.Tensor   # (B, n_p, 1) f32
    physics_param: torch.Tensor  # (B, 1) f32
    state_mask: torch.Tensor   # (B, N) bool
    eef_mask: torch.Tensor     # (B, N) bool
    obj_mask: torch.Tensor     # (B, n_p) bool
    adj_thresh: float
    topk: int
    connect_tools_all: bool
    n_p: int
    n_s: int

    @property
    def B(self):
        return self.state.shape[0]

    @property
    def N(self):
        return self.n_p + self.n_s

    def to(self, device):
        out = copy.copy(self)
        for k, v in vars(self).items():
            if torch.is_tensor(v):
                setattr(out, k, v.to(device))
        return out

    def take(self, sl) -> "Workload":
        """Sub-batch (slice or index tensor over graphs)."""
        out = copy.copy(self)
        for k, v in vars(self).items():
            if torch.is_tensor(v):
                setattr(out, k, v[sl])
        return out

    def graph_dict(self, Rr=None, Rs=None) -> Dict[str, torch.Tensor]:
        d = {
            "state": self.state, "action": self.action, "attrs": self.attrs,
            "p_instance": self.p_instance, "obj_mask": self.obj_mask,
            "state_mask": self.state_mask, "eef_mask": self.eef_mask,
            "p_rigid": torch.zeros(self.B, 1, device=self.state.device),
            f"{self.material}_physics_param": self.physics_param,
        }
        if Rr is not None:
            d["Rr"], d["Rs"] = Rr, Rs
        return d


def _object_layout(material: str, n_p: int, g: torch.Generator) -> torch.Tensor:
    i = torch.arange(n_p, dtype=torch.float32)
    if material == "rope":
        x = 0.2 * i
        return torch.stack([x, torch.zeros(n_p), 0.3 * torch.sin(x)], 1)
    side = int(math.ceil(math.sqrt(n_p)))
    gx, gz = (i % side), torch.div(i, side, rounding_mode="floor")
    if material == "granular":
        jit = (torch.rand(n_p, 2, generator=g) - 0.5) * 0.06
        y = torch.rand(n_p, generator=g) * 0.02
        return torch.stack([0.2 * gx + jit[:, 0], y, 0.2 * gz + jit[:, 1]], 1)
    if material == "cloth":
        return torch.stack([0.25 * gx, torch.zeros(n_p), 0.25 * gz], 1)
    raise ValueError(material)


def _tool_layout(material: str, obj: torch.Tensor, n_s: int) -> torch.Tensor:
    n_p = obj.shape[0]
    if material == "rope":
        base = obj[n_p // 2] + torch.tensor([0.05, 0.0, 0.15])
        return base[None].repeat(n_s, 1) + torch.arange(n_s)[:, None] * torch.tensor([0.1, 0.0, 0.0])
    if material == "granular":
        c = obj.mean(0)
        off = torch.linspace(-0.25, 0.25, n_s) if n_s > 1 else torch.zeros(1)
        return torch.stack([c[0] + off, torch.full((n_s,), 0.01), c[2].repeat(n_s)], 1)
    # cloth: grippers 0.05 above a corner
    off = torch.arange(n_s, dtype=torch.float32) * 0.1
    return torch.stack([obj[0, 0] + off, torch.full((n_s,), 0.05), obj[0, 2].repeat(n_s)], 1)


def make_workload(material: str, n_p: int, B: int, seed: int, n_his: int = 4,
                  n_s: int | None = None, n_pad: int = 0) -> Workload:
    """Build a batch of B graphs with n_p object particles (+ n_s tool particles).

    `n_pad` trailing object slots are marked invalid (state_mask False, zero
    position) the way DynDataset pads to max_nobj (dataset/dataset.py:170-176).
    """
    thr, topk, cta, n_s_def = MATERIALS[material]
    n_s = n_s_def if n_s is None else n_s
    g = torch.Generator().manual_seed(seed)
    N = n_p + n_s
    n_real = n_p - n_pad
    cur = torch.zeros(B, N, 3)
    for b in range(B):
        obj = _object_layout(material, n_real, g)
        tool = _tool_layout(material, obj, n_s)
        cur[b, :n_real] = obj + 0.01 * torch.randn(n_real, 3, generator=g)
        cur[b, n_p:] = tool + 0.01 * torch.randn(n_s, 3, generator=g)
    state = cur[:, None].repeat(1, n_his, 1, 1) + 0.002 * torch.randn(B, n_his, N, 3, generator=g)
    state[:, -1] = cur
    state[:, :, n_real:n_p] = 0.0
    attrs = torch.zeros(B, N, 2)
    attrs[:, :n_real, 0] = 1.0
    attrs[:, n_p:, 1] = 1.0
    action = torch.zeros(B, N, 3)
    action[:, n_p:] = torch.tensor([0.05, 0.0, 0.02])
    p_instance = torch.zeros(B, n_p, 1)
    p_instance[:, :n_real] = 1.0
    state_mask = torch.zeros(B, N, dtype=torch.bool)
    state_mask[:, :n_real] = True
    state_mask[:, n_p:] = True
    eef_mask = torch.zeros(B, N, dtype=torch.bool)
    eef_mask[:, n_p:] = True
    obj_mask = state_mask[:, :n_p].clone()
    return Workload(material, state, attrs, action, p_instance, torch.full((B, 1), 0.5),
                    state_mask, eef_mask, obj_mask, thr, topk, cta, n_p, n_s)


# BASELINE.json configs (SURVEY.md §8d "Config instances")
BASELINE_CONFIGS = {
    1: dict(material="rope", n_p=100, B=1, pstep=1, T=1),
    2: dict(material="rope", n_p=300, B=32, pstep=4, T=1),
    3: dict(material="granular", n_p=1000, B=64, pstep=3, T=5),
    4: dict(material="cloth", n_p=2000, B=128, pstep=3, T=10),
}


def baseline_workload(cfg_id: int, B: int | None = None) -> Workload:
    c = BASELINE_CONFIGS[cfg_id]
    return make_workload(c["material"], c["n_p"], c["B"] if B is None else B, seed=1234 + cfg_id)
