"""Relation-row padding helpers with the reference's signatures (src/dynamics/utils.py:37-46, :127-137).

They only move data (zero-fill + copy / slicing) and exist for callers that still carry dense
Rr / Rs; the sparse path (`EdgeList`) needs neither.
"""
import torch


def pad_torch(x, max_dim, dim=0):
    """Zero-pads `x` along `dim` (0 for (n_rel, N), 1 for (B, n_rel, N)) to `max_dim` rows.
    Like the reference it raises when x already has more rows than max_dim."""
    if dim == 0:
        out = torch.zeros((max_dim, x.shape[1]), dtype=x.dtype, device=x.device)
        out[:x.shape[0]] = x
    elif dim == 1:
        out = torch.zeros((x.shape[0], max_dim, x.shape[2]), dtype=x.dtype, device=x.device)
        out[:, :x.shape[1]] = x
    else:
        raise ValueError("pad_torch supports dim 0 or 1")
    return out


def truncate_graph(data):
    """Drops trailing all-zero relation rows of data['Rr'], data['Rs'] down to the largest per-graph count."""
    n_r = (data["Rr"].sum(-1) > 0).sum(1).max()
    n_s = (data["Rs"].sum(-1) > 0).sum(1).max()
    n = int(torch.maximum(n_r, n_s).item())
    data["Rr"] = data["Rr"][:, :n, :]
    data["Rs"] = data["Rs"][:, :n, :]
    return data
