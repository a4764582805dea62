"""Per-stage comparison of the CUDA forward against golden intermediates (GPU box diagnostic).

Runs agx_forward through ctypes with a workspace it keeps, then compares every internal buffer
(nfeat, P, A_n, Qr, Qs, agg, C_e) with values recomputed on the CPU from the golden per-stage tensors
of the reference.  Prints one line per buffer; exits non-zero if any differs by more than 1e-4.
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import agx_helpers as H  # noqa: E402


def main(fname="forward_cloth64_pad_k3.npz", precision=0):
    import adaptigraph_b200 as agx
    from adaptigraph_b200 import _lib as L, ops, synthetic as syn
    from oracle import dynamics_oracle as orc
    g = H.load_npz(fname)
    K = int(g["pstep"])
    mat = str(g["material"])
    p = H.golden_weights()
    m = agx.DynamicsPredictor(*syn.configs(mat, K), "cuda")
    m.load_state_dict(p)
    m = m.cuda().eval()
    t = lambda k: torch.from_numpy(g[k])  # noqa: E731
    state, attrs, action, p_inst, phys = t("state"), t("attrs"), t("action"), t("p_instance"), t("physics_param")
    B, Hh, N, _ = state.shape
    n_p = p_inst.shape[1]
    row_ptr, send = H.csr_from_lists(g["recv"], g["send"], N)
    deg = (row_ptr[1:] - row_ptr[:-1]).long()
    recv = torch.repeat_interleave(torch.arange(B * N), deg).to(torch.int32)
    E = int(row_ptr[-1])
    dims = ops.make_dims(150, Hh, 2, 1, 3, K)
    d = lambda x: x.cuda().contiguous()  # noqa: E731
    st, at, ac, pi, ph = d(state), d(attrs), d(action), d(p_inst[:, :, 0]), d(phys)
    rp, sd, rv = d(row_ptr), d(send), d(recv)
    gin = L.AgxGraphIn(B, N, n_p, st.data_ptr(), at.data_ptr(), ac.data_ptr(), pi.data_ptr(), ph.data_ptr(),
                       rp.data_ptr(), sd.data_ptr(), rv.data_ptr(), E)
    nws = L.lib.agx_forward_workspace_bytes(C.byref(dims), B, N, E)
    ws = torch.zeros(nws, dtype=torch.uint8, device="cuda")
    pos = torch.empty(B, n_p, 3, device="cuda")
    mot = torch.empty(B, n_p, 3, device="cuda")
    packed = m.packed_weights()
    L.check(L.lib.agx_forward(C.byref(dims), C.c_void_p(packed.data_ptr()), C.byref(gin), C.c_void_p(pos.data_ptr()), n_p * 3,
                              C.c_void_p(mot.data_ptr()), precision, C.c_void_p(ws.data_ptr()), nws, None), "agx_forward")
    torch.cuda.synchronize()
    wsf = ws.cpu().view(torch.float32)
    rows = B * N
    off = 0

    def take(n):
        nonlocal off
        off = (off + 63) // 64 * 64          # 256-byte alignment in floats
        out = wsf[off:off + n]
        off += n
        return out
    nfeat = take(rows * 16).view(rows, 16)
    pad = lambda n: (n + 127) // 128 * 128      # the matrices are padded to whole 128-row tiles

    def matrix(n):
        """[n][160] matrix: row-major in the fp32 path, [tile][piece of 16 columns][row in tile][16] in the tensor-core path."""
        raw = take(pad(n) * 160)
        if precision == 1:
            return raw.view(pad(n) // 128, 10, 128, 16).permute(0, 2, 1, 3).reshape(pad(n), 160)[:n]
        return raw.view(pad(n), 160)[:n]
    P = matrix(rows)[:, :150]
    A = matrix(rows)[:, :150]
    Qr = matrix(rows)[:, :150]
    Qs = matrix(rows)[:, :150]
    agg_raw = matrix(rows)
    Cb = matrix(max(E, 1))[:E, :150]
    take(rows); take(rows)                      # rowmaxP, rowmaxA
    agg_exp = take(rows).view(torch.int32)
    if precision == 1:                          # tensor-core path: every 16-word piece holds 8 packed fp16 hi pairs then 8 lo pairs
        words = agg_raw.contiguous().view(rows, 10, 16)
        hi = words[:, :, :8].contiguous().view(torch.float16).view(rows, 160).float()
        lo = words[:, :, 8:].contiguous().view(torch.float16).view(rows, 160).float()
        aggb = ((hi + lo) * torch.exp2(-agg_exp.float())[:, None])[:, :150]
    else:
        aggb = agg_raw[:, :150]

    # expectations from golden per-stage tensors
    hist, p_in, group = orc.node_and_relation_inputs(state, attrs, p_inst, action, phys)
    nfeat_ref = torch.cat([hist.reshape(rows, 12), attrs.reshape(rows, 2), group.reshape(rows, 1), torch.zeros(rows, 1)], 1)
    penc = torch.from_numpy(g["particle_encode"]).reshape(rows, 150)
    renc_dense = torch.from_numpy(g["relation_encode"])                       # (B, n_rel, 150), padded rows included
    keep = torch.from_numpy(g["recv"] >= 0)
    renc = renc_dense[keep]                                                    # reference row order == CSR order
    eff = torch.from_numpy(g["particle_effect"]).reshape(K, rows, 150)
    Wr, br = p["relation_propagator.linear.weight"], p["relation_propagator.linear.bias"]
    Wp, bp = p["particle_propagator.linear.weight"], p["particle_propagator.linear.bias"]
    C_ref = renc @ Wr[:, :150].T + br
    A_ref = penc @ Wp[:, :150].T + bp
    P_last_written = penc if K == 1 else eff[K - 2]
    Qr_ref, Qs_ref = P_last_written @ Wr[:, 150:300].T, P_last_written @ Wr[:, 300:].T
    snd = send.long() + (recv.long() // N) * N
    e_out = torch.relu(C_ref + Qr_ref[recv.long()] + Qs_ref[snd])
    agg_ref = torch.zeros(rows, 150).index_add_(0, recv.long(), e_out)
    bad = 0
    for name, got, ref in [("nfeat", nfeat, nfeat_ref), ("A_n", A, A_ref), ("C_e", Cb, C_ref), ("P(last written)", P, P_last_written),
                           ("Qr", Qr, Qr_ref), ("Qs", Qs, Qs_ref), ("agg(last pstep)", aggb, agg_ref),
                           ("pred_motion", mot.cpu(), t("pred_motion")), ("pred_pos", pos.cpu(), t("pred_pos"))]:
        err = (got - ref).abs().max().item() if ref.numel() else 0.0
        flag = "OK " if err <= 1e-4 else "BAD"
        bad += flag == "BAD"
        print(f"prec={precision} {flag} {name:18s} max-abs {err:.3e}  ref-rms {ref.pow(2).mean().sqrt().item() if ref.numel() else 0:.3e}")
    return bad


if __name__ == "__main__":
    prec = int(os.environ.get("AGX_DEBUG_PREC", "0"))
    files = [x for x in sys.argv[1:]] or ["forward_rope100_k1.npz", "forward_cloth64_pad_k3.npz"]
    sys.exit(1 if sum(main(f, prec) for f in files) else 0)
