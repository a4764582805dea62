"""Shared test helpers: golden loading and edge-format conversion (CPU only)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_npz(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: z[k] for k in z.files}


def golden_weights():
    return {k: torch.from_numpy(v) for k, v in load_npz("weights_seed0.npz").items()}


def onehots_from_lists(recv, send, N):
    """int (B, n_rel) receiver/sender ids with -1 padding -> dense f32 (B, n_rel, N) Rr, Rs."""
    recv, send = torch.as_tensor(recv).long(), torch.as_tensor(send).long()
    B, R = recv.shape
    Rr, Rs = torch.zeros(B, R, N), torch.zeros(B, R, N)
    b, e = torch.nonzero(recv >= 0, as_tuple=True)
    Rr[b, e, recv[b, e]] = 1
    Rs[b, e, send[b, e]] = 1
    return Rr, Rs


def lists_from_onehots(Rr, Rs):
    def one(R):
        idx = R.argmax(-1).to(torch.int32)
        idx[R.sum(-1) == 0] = -1
        return idx
    return one(Rr), one(Rs)


def csr_from_lists(recv, send, N):
    """-1 padded (B, n_rel) lists (rows sorted by receiver, as the reference emits) -> CSR over B*N rows."""
    recv, send = torch.as_tensor(recv).long(), torch.as_tensor(send).long()
    B = recv.shape[0]
    deg = torch.zeros(B * N, dtype=torch.long)
    sends = []
    for b in range(B):
        ok = recv[b] >= 0
        r, s = recv[b][ok], send[b][ok]
        assert bool((r[1:] >= r[:-1]).all()), "relation rows must be receiver-sorted"
        deg.index_add_(0, r + b * N, torch.ones_like(r))
        sends.append(s)
    row_ptr = torch.zeros(B * N + 1, dtype=torch.int32)
    row_ptr[1:] = torch.cumsum(deg, 0).to(torch.int32)
    return row_ptr, torch.cat(sends).to(torch.int32)


def graph_case_names(cases):
    return sorted({k.split("/")[0] for k in cases})
