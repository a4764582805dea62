"""Batch sharding across GPUs: graphs are independent (model.py:134-313 has no cross-batch term), so every rank
runs graph build + forward/rollout on a contiguous slice of the batch and no data-path collective is needed.
The helpers below are host-side plumbing over torch.distributed (NCCL on GPUs, gloo in CPU tests)."""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist


def shard_slice(B: int, world: int, rank: int) -> slice:
    """Contiguous split of B graphs over `world` ranks; the first B % world ranks get one extra graph."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return slice(lo, lo + base + (1 if rank < rem else 0))


def shard_graph_dict(graph: Dict[str, torch.Tensor], world: int, rank: int) -> Dict[str, torch.Tensor]:
    """Slices every batched tensor of a reference-style graph dict (forward_dynamics.py:130-147)."""
    B = graph["state"].shape[0]
    sl = shard_slice(B, world, rank)
    return {k: (v[sl] if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == B else v) for k, v in graph.items()}


def gather_batch(local: torch.Tensor, B: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """All-gathers per-rank shards (possibly uneven) back into batch order; only needed when a caller wants the
    full prediction on every rank (MPPI needs one reward per sample, `planner.py:249`)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_slice(B, world, r) for r in range(world)]
    width = max(s.stop - s.start for s in sizes)
    pad = local.new_zeros((width,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[: s.stop - s.start] for p, s in zip(parts, sizes)], 0)


def max_over_ranks(value: float, device, group: Optional[dist.ProcessGroup] = None) -> float:
    """Timing rule: multi-GPU numbers are the max over ranks of a device-side measurement."""
    if not (dist.is_available() and dist.is_initialized()):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def allreduce_gradients(parameters, group: Optional[dist.ProcessGroup] = None, average: bool = True) -> None:
    """Data-parallel training: ONE all-reduce of a flat fp32 bucket holding every gradient
    (252,903 floats = 1.01 MB for the shipped model), then scatter back.  NCCL over NVLink on GPUs, gloo in CPU tests."""
    grads = [p.grad for p in parameters if p.grad is not None]
    if not grads or not (dist.is_available() and dist.is_initialized()):
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
