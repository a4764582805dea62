"""ORACLE — TEST INFRASTRUCTURE ONLY.  Compiles oracle/graph_oracle.c into oracle/libgraph_oracle.so
and exposes it through ctypes.  (The reference is pure Python, so there is no oracle/_ref binary:
the reference itself was run in the build container to produce tests/golden/*.npz.)"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "graph_oracle.c")
LIB = os.path.join(HERE, "libgraph_oracle.so")


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", SRC, "-o", LIB], check=True)
    return LIB


def edges(pos, mask, tool, thr2, topk, cta, sem=0):
    """numpy in, numpy out: (recv, send, n_edges) with graph-local ids, reference row order."""
    lib = C.CDLL(build())
    lib.agx_oracle_edges.restype = C.c_int64
    pos = np.ascontiguousarray(pos, np.float32)
    B, N, _ = pos.shape
    mask = np.ascontiguousarray(mask, np.uint8)
    tool = np.ascontiguousarray(tool, np.uint8)
    thr2 = np.ascontiguousarray(np.broadcast_to(np.asarray(thr2, np.float32), (B,)))
    cap = B * N * (min(topk, N) + int(tool.sum(-1).max()) + 1)
    recv = np.empty(cap, np.int32)
    send = np.empty(cap, np.int32)
    n_edges = np.empty(B, np.int32)
    vp = C.c_void_p
    tot = lib.agx_oracle_edges(vp(pos.ctypes.data), vp(mask.ctypes.data), vp(tool.ctypes.data), vp(thr2.ctypes.data),
                               C.c_int32(B), C.c_int32(N), C.c_int32(topk), C.c_int32(int(cta)), C.c_int32(sem),
                               vp(recv.ctypes.data), vp(send.ctypes.data), C.c_int64(cap), vp(n_edges.ctypes.data))
    assert tot >= 0, tot
    return recv[:tot].copy(), send[:tot].copy(), n_edges


if __name__ == "__main__":
    print(build(force=True))
