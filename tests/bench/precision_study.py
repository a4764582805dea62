"""CPU study of the tensor-core path's precision budget (SURVEY.md §7 hard part 2).

Emulates, in float64 plus explicit roundings, the arithmetic variants the CUDA chains could run and measures the 10-step
rollout RMSE of each against the exact (float64) forward on the same inputs, with the relations rebuilt every step
(forward_dynamics.py:156-197):

  * per dense layer: how many of the split-fp16 partial products are issued
        3  A(22 bit) x W(22 bit)      Alo*Whi + Ahi*Wlo + Ahi*Whi   (current default)
        2a A(22 bit) x W(11 bit)      Alo*Whi + Ahi*Whi             (no lo weight image: half the shared memory)
        2b A(11 bit) x W(22 bit)      Ahi*Wlo + Ahi*Whi
        15 as 2b with the Wlo term in fp8 (e5m2 activations x e4m3 weight residuals): 1.5 MMA units
        1  A(11 bit) x W(11 bit)      Ahi*Whi
        8  main term in fp16, the two correction terms in fp8 e4m3 (half the tensor time each)
  * storage of the write-once / read-many streams (C per relation, Qr / Qs per particle):
        f32, or 16-bit fixed point with one power-of-two scale per (row, 16-column piece) ("i16"), or fp16 with a row scale

Usage: python tests/bench/precision_study.py [--material cloth --n_p 2000 --B 2 --T 10]
Test infrastructure (imports oracle/); prints one JSON line per variant.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import dynamics_oracle as orc  # noqa: E402
from adaptigraph_b200 import synthetic as syn  # noqa: E402

TARGET_EXP = 14


def pow2_scale(bound):
    """2^e with bound * 2^e <= 2^14 (the kernels' scale_exp)."""
    b = bound.clamp_min(1e-30)
    e = TARGET_EXP - torch.ceil(torch.log2(b))
    return torch.pow(2.0, e.clamp(-100, 100))


def f16(x):
    return x.to(torch.float32).to(torch.float16).to(torch.float64)


def e4m3(x):
    return x.to(torch.float32).clamp(-448, 448).to(torch.float8_e4m3fn).to(torch.float32).to(torch.float64)


def split(x, s):
    xs = x * s
    hi = f16(xs)
    lo = f16(xs - hi)
    return hi, lo


def qlinear(x, W, b, mode):
    """y = x W^T + b with the products of `mode`; x (M,K) float64 (already the fp32 values), scales exact powers of two."""
    if mode == "exact":
        return x @ W.T + b
    sa = pow2_scale(x.abs().amax(1, keepdim=True).clamp_min(1.0))
    sw = pow2_scale(torch.maximum(W.abs().max(), b.abs().max()))
    ahi, alo = split(x, sa)
    whi, wlo = split(W, sw)
    bhi, blo = split(b, sw)
    one = sa                      # the constant-1 column carries the row scale
    if mode == "3":
        acc = (ahi + alo) @ whi.T + ahi @ wlo.T + one * (bhi + blo)
    elif mode == "2a":
        acc = (ahi + alo) @ whi.T + one * (bhi + blo)      # bias lo rides in a second padding column of the hi image
    elif mode == "2b":
        acc = ahi @ (whi + wlo).T + one * (bhi + blo)
    elif mode == "1":
        acc = ahi @ whi.T + one * (bhi + blo)
    elif mode == "15":     # fp16 main term + ONE fp8 correction: e5m2(A) x e4m3(Wlo), half the tensor time of an fp16 MMA
        a8 = (ahi * 2.0 ** -6).to(torch.float32).to(torch.float8_e5m2).to(torch.float32).to(torch.float64)
        w8 = e4m3(wlo * 2.0 ** 6)
        acc = ahi @ whi.T + a8 @ w8.T + one * (bhi + blo)
    elif mode == "15u":    # the same without the 2^6 shifts (saves a multiply per activation in the epilogue)
        a8 = ahi.to(torch.float32).to(torch.float8_e5m2).to(torch.float32).to(torch.float64)
        acc = ahi @ whi.T + a8 @ e4m3(wlo).T + one * (bhi + blo)
    elif mode == "8":
        # corrections in fp8: (Alo * 2^u) x (Whi * 2^-u) etc. with shifts that keep both operands inside e4m3's range
        a8lo = e4m3(alo * 2.0 ** 4)            # |alo| <= 2^3  -> 2^7
        w8hi = e4m3(whi * 2.0 ** -6)           # |whi| <= 2^14 -> 2^8
        a8hi = e4m3(ahi * 2.0 ** -6)
        w8lo = e4m3(wlo * 2.0 ** 4)
        acc = ahi @ whi.T + (a8lo @ w8hi.T) * 2.0 ** 2 + (a8hi @ w8lo.T) * 2.0 ** 2 + one * (bhi + blo)
    else:
        raise ValueError(mode)
    return acc / (sa * sw)


def store(x, fmt):
    """Round-trip of a stored [rows][150] matrix."""
    if fmt == "f32":
        return x.to(torch.float32).to(torch.float64)
    if fmt == "i16":       # 16-bit fixed point, one power-of-two scale per (row, 16-column piece)
        R, F = x.shape
        pad = (-F) % 16
        xp = torch.nn.functional.pad(x, (0, pad)).reshape(R, -1, 16)
        mx = xp.abs().amax(2, keepdim=True).clamp_min(1e-30)
        s = torch.pow(2.0, 14 - torch.ceil(torch.log2(mx)))          # |x| * s <= 2^14 < 32767
        q = torch.round(xp * s).clamp(-32767, 32767) / s
        return q.reshape(R, -1)[:, :F]
    if fmt == "i16row":    # one scale per row
        mx = x.abs().amax(1, keepdim=True).clamp_min(1e-30)
        s = torch.pow(2.0, 14 - torch.ceil(torch.log2(mx)))
        return torch.round(x * s) / s
    if fmt == "f16row":
        mx = x.abs().amax(1, keepdim=True).clamp_min(1e-30)
        s = pow2_scale(mx)
        return f16(x * s) / s
    raise ValueError(fmt)


def forward_variant(p, pstep, state, attrs, row_ptr, send, p_instance, action, physics_param, cfg):
    """oracle.forward_sparse (model.py:129-313 with the hoist) in float64 with the roundings of `cfg`."""
    B, H, N, _ = state.shape
    n_p = p_instance.shape[1]
    D = torch.float64
    hist, p_in, group = orc.node_and_relation_inputs(state, attrs, p_instance, action, physics_param)
    deg = (row_ptr[1:] - row_ptr[:-1]).long()
    recv = torch.repeat_interleave(torch.arange(B * N), deg)
    snd = send.long() + (recv // N) * N
    fl = lambda t: t.reshape(B * N, -1).to(D)  # noqa: E731
    a, g, hs = fl(attrs), fl(group), fl(hist)
    rel_in = torch.cat([a[recv], a[snd], (g[recv] - g[snd]).abs().sum(1, keepdim=True), hs[recv] - hs[snd]], 1)
    rel_in = rel_in.to(torch.float32).to(D)      # the kernels form the differences in fp32
    F = 150
    W = lambda n: p[n + ".weight"].to(D)  # noqa: E731
    bb = lambda n: p[n + ".bias"].to(D)  # noqa: E731
    m = cfg["mma"]
    relu = torch.relu
    x = fl(p_in)
    for i, key in zip((0, 2, 4), ("penc0", "penc2", "penc4")):
        x = relu(qlinear(x, W(f"particle_encoder.model.{i}"), bb(f"particle_encoder.model.{i}"), m.get(key, "3")))
    penc = x
    x = rel_in
    for i, key in zip((0, 2, 4), ("renc0", "renc2", "renc4")):
        x = relu(qlinear(x, W(f"relation_encoder.model.{i}"), bb(f"relation_encoder.model.{i}"), m.get(key, "3")))
    renc = x
    Wr, br = W("relation_propagator.linear"), bb("relation_propagator.linear")
    Wp, bp = W("particle_propagator.linear"), bb("particle_propagator.linear")
    zero = torch.zeros(F, dtype=D)
    c_edge = store(qlinear(renc, Wr[:, :F], br, m.get("rp_rel", "3")), cfg.get("C", "f32"))
    a_node = store(qlinear(penc, Wp[:, :F], bp, m.get("pp_enc", "3")), cfg.get("A", "f32"))
    eff = store(penc, "f32")
    for _ in range(pstep):
        q_r = store(qlinear(eff, Wr[:, F:2 * F], zero, m.get("rp_recv", "3")), cfg.get("Qr", cfg.get("Q", "f32")))
        q_s = store(qlinear(eff, Wr[:, 2 * F:], zero, m.get("rp_send", "3")), cfg.get("Qs", cfg.get("Q", "f32")))
        e_out = relu(c_edge + q_r[recv] + q_s[snd])
        agg = store(torch.zeros_like(eff).index_add_(0, recv, e_out), cfg.get("agg", "f32"))
        eff = store(relu(a_node + qlinear(agg, Wp[:, F:], zero, m.get("pp_agg", "3")) + eff), "f32")
    eff = eff.reshape(B, N, F)[:, :n_p].reshape(-1, F)
    h = relu(qlinear(eff, W("non_rigid_predictor.linear_0"), bb("non_rigid_predictor.linear_0"), m.get("pred0", "3")))
    h = relu(qlinear(h, W("non_rigid_predictor.linear_1"), bb("non_rigid_predictor.linear_1"), m.get("pred1", "3")))
    motion = (h @ W("non_rigid_predictor.linear_2").T + bb("non_rigid_predictor.linear_2")).reshape(B, n_p, 3)
    return (state[:, -1, :n_p].to(D) + motion.clamp(-100, 100)).to(torch.float32), motion


def rollout_variant(p, pstep, w, T, cfg, frozen=None):
    state = w.state.clone()
    n_p = w.n_p
    out, edges = [], []
    for t in range(T):
        if frozen is not None:
            row_ptr, send = frozen[t]
        else:
            adj = orc.adjacency_batch(state[:, -1], w.adj_thresh, w.state_mask, w.eef_mask, w.topk, w.connect_tools_all)
            row_ptr, send = orc.edge_lists_from_adjacency(adj)
        edges.append((row_ptr, send))
        pred, _ = forward_variant(p, pstep, state, w.attrs, row_ptr, send, w.p_instance, w.action, w.physics_param, cfg)
        out.append(pred)
        y = pred[:, :, 1].min(1).values
        tool = state[:, -1, n_p:] + w.action[:, n_p:]
        tool[:, :, 1] = y[:, None]
        cur = torch.cat([pred, tool], 1)
        state = torch.cat([state[:, 1:], cur[:, None]], 1)
    return torch.stack(out, 1), edges


EDGE = ("renc2", "renc4", "rp_rel")
NODE_ENC = ("penc2", "penc4", "pp_enc", "rp_recv", "rp_send")
UPD = ("pp_agg", "pred0", "pred1")


def variants():
    v = {"3-everywhere": dict(mma={})}
    for mode in ("2a", "2b", "15", "15u", "1", "8"):
        v[f"edge-chain {mode}"] = dict(mma={k: mode for k in EDGE + ("renc0",)})
        v[f"all layers {mode}"] = dict(mma={k: mode for k in EDGE + NODE_ENC + UPD + ("renc0", "penc0")})
    for k in EDGE:
        v[f"only {k} 2a"] = dict(mma={k: "2a"})
        v[f"only {k} 1"] = dict(mma={k: "1"})
    v["renc2+renc4 2a, rp_rel 3"] = dict(mma={"renc2": "2a", "renc4": "2a"})
    v["renc2+renc4 1, rp_rel 3"] = dict(mma={"renc2": "1", "renc4": "1"})
    v["renc2+renc4 1, rp_rel 2a"] = dict(mma={"renc2": "1", "renc4": "1", "rp_rel": "2a"})
    for fmt in ("i16", "i16row", "f16row"):
        v[f"C {fmt}"] = dict(mma={}, C=fmt)
        v[f"C+Q {fmt}"] = dict(mma={}, C=fmt, Q=fmt)
    v["edge-chain 2a + C,Q i16"] = dict(mma={k: "2a" for k in EDGE + ("renc0",)}, C="i16", Q="i16")
    v["edge-chain 8 + C,Q i16"] = dict(mma={k: "8" for k in EDGE + ("renc0",)}, C="i16", Q="i16")
    for k in NODE_ENC + UPD + ("renc0", "penc0"):
        v[f"only {k} 2b"] = dict(mma={k: "2b"})
    v["node-encoder chain 2b"] = dict(mma={k: "2b" for k in NODE_ENC + ("penc0",)})
    v["update chains 2b"] = dict(mma={k: "2b" for k in UPD + ("rp_recv", "rp_send")})
    v["edge-chain 2b + C i16"] = dict(mma={k: "2b" for k in EDGE + ("renc0",)}, C="i16")
    v["edge-chain 2b + penc 2b + C i16"] = dict(mma={k: "2b" for k in EDGE + ("renc0", "penc0", "penc2", "penc4")}, C="i16")
    v["edge-chain 2b + C i16 + A i16"] = dict(mma={k: "2b" for k in EDGE + ("renc0",)}, C="i16", A="i16")
    v["edge-chain 2b + C i16 + agg i16"] = dict(mma={k: "2b" for k in EDGE + ("renc0",)}, C="i16", agg="i16")
    v["edge-chain 2b + C i16 + A,agg i16"] = dict(mma={k: "2b" for k in EDGE + ("renc0",)}, C="i16", A="i16", agg="i16")
    v["edge-chain 2b + C i16 + A,agg,Q i16"] = dict(mma={k: "2b" for k in EDGE + ("renc0",)}, C="i16", A="i16", agg="i16", Q="i16")
    v["shipped (edge-chain 2b + C i16) + Qs i16"] = dict(mma={k: "2b" for k in EDGE + ("renc0",)}, C="i16", Qs="i16")
    v["shipped (edge-chain 2b + C i16) + Qs,Qr i16"] = dict(mma={k: "2b" for k in EDGE + ("renc0",)}, C="i16", Q="i16")
    v["only Qs i16"] = dict(mma={}, Qs="i16")
    v["only Qr i16"] = dict(mma={}, Qr="i16")
    v["only A i16"] = dict(mma={}, A="i16")
    v["only agg i16"] = dict(mma={}, agg="i16")
    v["all 2a + C,Q i16"] = dict(mma={k: "2a" for k in EDGE + NODE_ENC + UPD + ("renc0", "penc0")}, C="i16", Q="i16")
    return v


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--material", default="cloth")
    ap.add_argument("--n_p", type=int, default=2000)
    ap.add_argument("--B", type=int, default=2)
    ap.add_argument("--T", type=int, default=10)
    ap.add_argument("--pstep", type=int, default=3)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    w = syn.make_workload(args.material, args.n_p, args.B, seed=1238)
    # the reference constructor's weights under manual_seed(0) (tests/golden/weights_seed0.npz)
    import numpy as np
    gold = np.load(os.path.join(ROOT, "tests", "golden", "weights_seed0.npz"))
    p = {k: torch.from_numpy(gold[k]) for k in gold.files}
    exact, edges = rollout_variant(p, args.pstep, w, args.T, dict(mma={k: "exact" for k in EDGE + NODE_ENC + UPD + ("renc0", "penc0")}))
    f32, _ = orc.rollout_dense(p, args.pstep, w.state, w.attrs, w.p_instance, w.action, w.physics_param, w.state_mask, w.eef_mask,
                               w.adj_thresh, w.topk, w.connect_tools_all, args.T) if args.n_p <= 1000 else (None, None)
    if f32 is not None:
        e = (f32 - exact).double()
        print(json.dumps({"variant": "torch fp32 dense (the reference arithmetic)", "rmse": float(e.pow(2).mean().sqrt()), "max": float(e.abs().max())}), flush=True)
    motion = (exact[:, 1:] - exact[:, :-1]).double().pow(2).mean().sqrt()
    print(json.dumps({"info": "rms displacement per step", "value": float(motion)}), flush=True)
    for name, cfg in variants().items():
        if args.only and args.only not in name:
            continue
        got, ed = rollout_variant(p, args.pstep, w, args.T, cfg)
        e = (got - exact).double()
        same = all(torch.equal(a[1], b[1]) and torch.equal(a[0], b[0]) for a, b in zip(ed, edges))
        got_f, _ = rollout_variant(p, args.pstep, w, args.T, cfg, frozen=edges)
        ef = (got_f - exact).double()
        print(json.dumps({"variant": name, "rmse": float(e.pow(2).mean().sqrt()), "max": float(e.abs().max()),
                          "rmse_step1": float(e[:, 0].pow(2).mean().sqrt()), "same_relations": same,
                          "rmse_frozen_relations": float(ef.pow(2).mean().sqrt())}), flush=True)


if __name__ == "__main__":
    main()
