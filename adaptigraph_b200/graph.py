"""Relation construction: drop-in `construct_edges_from_states[_batch]` and the sparse fast path.

Reference: src/dynamics/dataset/graph.py:38-89 (single graph) and :91-156 (batched).
The reference returns two dense one-hot matrices Rr, Rs of shape (B, n_rel, N); the
engine works on CSR edge lists (`EdgeList`).  `construct_edges_from_states[_batch]`
keep the reference's signatures and dense return type (rows in the reference's order,
n_rel = max over the batch, zero-padded) for callers that have not switched;
`build_edges` returns the `EdgeList` that `DynamicsPredictor.forward(..., edges=...)`
and `rollout` consume without ever materialising the one-hots.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple, Union

import numpy as np
import torch
from torch import Tensor

from . import _lib as L
from . import ops


@dataclass
class EdgeList:
    """CSR-by-receiver relations of a batch of graphs (layout: include/adaptigraph_b200.h)."""
    row_ptr: Tensor   # int32 (B*N+1)
    send: Tensor      # int32 (cap) graph-local sender ids, valid in [0, row_ptr[-1])
    recv: Tensor      # int32 (cap) flattened receiver ids
    n_edges: Tensor   # int32 (B) relations per graph
    status: Tensor    # int32 (1) bit0: capacity overflow
    B: int
    N: int

    def check(self) -> "EdgeList":
        """Synchronises and raises like pad_torch (utils.py:37-46) if the capacity was exceeded."""
        if int(self.status.item()) & 1:
            raise RuntimeError(f"relation capacity {self.send.numel()} exceeded: the batch has "
                               f"{int(self.row_ptr[-1].item())} relations (raise max_nR)")
        return self

    def to_dense(self, n_rel: Optional[int] = None) -> Tuple[Tensor, Tensor]:
        """Dense (B, n_rel, N) one-hots; n_rel defaults to the largest per-graph count (graph.py:146-147)."""
        if n_rel is None:
            n_rel = int(self.n_edges.max().item())
        return ops.edges_to_onehot(self.row_ptr, self.send, self.B, self.N, n_rel)


def _thr2_batch(adj_thresh: Union[float, Tensor], B: int, device) -> Tensor:
    # graph.py:106-108: a float becomes an fp32 tensor and is squared in fp32
    if isinstance(adj_thresh, (float, int)):
        t = np.float32(adj_thresh)
        return torch.full((B,), float(t * t), dtype=torch.float32, device=device)
    t = adj_thresh.to(device=device, dtype=torch.float32).reshape(-1)
    if t.numel() == 1:
        t = t.repeat(B)
    return t * t


def build_edges(states: Tensor, adj_thresh: Union[float, Tensor], mask: Tensor, tool_mask: Tensor, topk: int = 10,
                connect_tools_all: bool = False, max_nR: Optional[int] = None,
                semantics: int = L.AGX_SEM_BATCH) -> EdgeList:
    """Batched radius AND top-k relations as an EdgeList.  `max_nR` (relations per graph, as in the
    reference configs) bounds the output buffers; None sizes them for the worst case."""
    B, N, _ = states.shape
    k = min(topk, N)
    if max_nR is None:
        n_tool = int(tool_mask.sum(-1).max().item()) if connect_tools_all else 0
        cap = B * N * (k + n_tool)
    else:
        cap = B * max_nR
    thr2 = _thr2_batch(adj_thresh, B, states.device)
    row_ptr, send, recv, n_edges, status = ops.graph_build(states, mask, tool_mask, thr2, topk, connect_tools_all,
                                                           semantics, cap)
    return EdgeList(row_ptr, send, recv, n_edges, status, B, N)


def construct_edges_from_states_batch(states, adj_thresh, mask, tool_mask, topk=10, connect_tools_all=False):
    """Drop-in for graph.py:91-156: (B,N,3) states -> dense Rr, Rs (B, n_rel, N)."""
    return build_edges(states, adj_thresh, mask, tool_mask, topk, connect_tools_all).to_dense()


def construct_edges_from_states(states, adj_thresh, mask, tool_mask, topk=10, connect_tools_all=False):
    """Drop-in for graph.py:38-89: (N,3) states -> dense Rr, Rs (n_rel, N).  Note the single-graph
    builder squares the threshold in Python floats (:53) and applies connect_tools_all
    unconditionally, clearing tool-tool pairs (:77-80)."""
    N = states.shape[0]
    thr2 = torch.full((1,), float(np.float32(float(adj_thresh) * float(adj_thresh))), dtype=torch.float32,
                      device=states.device)
    n_tool = int(tool_mask.sum().item()) if connect_tools_all else 0
    cap = N * (min(topk, N) + n_tool)
    row_ptr, send, recv, n_edges, status = ops.graph_build(states[None], mask[None], tool_mask[None], thr2, topk,
                                                           connect_tools_all, L.AGX_SEM_SINGLE, cap)
    Rr, Rs = EdgeList(row_ptr, send, recv, n_edges, status, 1, N).to_dense()
    return Rr[0], Rs[0]


def edges_from_onehots(Rr: Tensor, Rs: Tensor) -> EdgeList:
    """Dense (B, n_rel, N) Rr/Rs (any row order, zero rows allowed anywhere) -> EdgeList.

    The compatibility path of DynamicsPredictor.forward: per-row ids come from the library
    (agx_onehot_to_ids); ordering rows by receiver is a stable device sort, so relations of one
    receiver keep the caller's relative order.  No host synchronisation.
    """
    N = Rr.shape[2]
    return collate_relation_lists(ops.onehot_to_ids(Rr), ops.onehot_to_ids(Rs), N, Rr.device)


# ------------------------------------------------------------------------------------------ sparse sample format (data loading)
def relation_lists(Rr, Rs):
    """One sample's dense (n_rel, N) one-hots (CPU or CUDA, zero rows = padding, dataset.py:215-219) -> two int32 vectors
    (receiver id, sender id per relation row, -1 on padding rows).  A pure format conversion (arg-max of one-hot rows): what a
    DataLoader worker ships instead of 2 x max_nR x N floats (~0.8 MB -> ~4 kB per sample for the shipped configs)."""
    def one(R):
        idx = R.argmax(-1).to(torch.int32)
        idx[R.sum(-1) == 0] = -1
        return idx
    return one(Rr), one(Rs)


def collate_relation_lists(recv: Tensor, send: Tensor, N: int, device=None) -> EdgeList:
    """Batched (B, n_rel) receiver / sender id lists (-1 = padding, any row order) -> EdgeList on `device` (default: current CUDA
    device).  Same stable receiver sort as `edges_from_onehots`, so relations of one receiver keep the sample's order."""
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    rid, sid = recv.to(dev).long(), send.to(dev).long()
    B = rid.shape[0]
    valid = (rid >= 0) & (sid >= 0)
    key = torch.where(valid, rid + torch.arange(B, device=dev)[:, None] * N, torch.full_like(rid, B * N)).reshape(-1)
    key_sorted, perm = torch.sort(key, stable=True)
    deg = torch.zeros(B * N + 1, dtype=torch.int64, device=dev).index_add_(0, key_sorted, torch.ones_like(key_sorted))[: B * N]
    row_ptr = torch.zeros(B * N + 1, dtype=torch.int32, device=dev)
    row_ptr[1:] = torch.cumsum(deg, 0).to(torch.int32)
    send_out = sid.reshape(-1)[perm].clamp_(min=0).to(torch.int32).contiguous()
    recv_out = key_sorted.clamp_(max=B * N - 1).to(torch.int32).contiguous()
    return EdgeList(row_ptr, send_out, recv_out, valid.sum(1).to(torch.int32), torch.zeros(1, dtype=torch.int32, device=dev), B, N)
