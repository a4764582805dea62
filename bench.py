"""bench.py — particle-steps/s of the particle-graph dynamics hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl engine|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[3], the one the north-star target is quoted on and
that fits one GPU — cloth, 2000 object particles + 2 tool particles per graph, batch 128 graphs per GPU,
pstep 3, 10-step autoregressive rollout with the relations rebuilt every step.
One bench "step" = one such rollout over the batch; particle-steps = B * n_p * T.  With N GPUs every
rank rolls out its own 128 graphs (weak scaling, no data-path collective).

Engine arm (default) prints one JSON line with
  value    device-resident throughput (inputs in HBM, CUDA events, max over ranks)
  e2e      same rollout through the public API from pinned HOST buffers, H2D + D2H inside the timed region
  roofline the dominant kernel's achieved algorithmic FLOP/s or B/s against MEASURED_PEAKS.json
  cpu_baseline the oracle's dense restatement of the reference on the host cores, bounded sample
Reference arm (--impl reference) times that same CPU path alone (the reference is pure Python and
does not travel to the GPU box; the oracle port is pinned to it by tests/golden).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOAD = dict(material="cloth", n_p=2000, B=128, pstep=3, T=10, max_nR=16384)
F = 150


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"], bf16_tflops_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------ CPU path
def cpu_rollout(w, params, pstep, T):
    """The reference's CPU path (dense one-hot relations, torch fp32) restated by the oracle."""
    from oracle import dynamics_oracle as orc
    preds, _ = orc.rollout_dense(params, pstep, w.state, w.attrs, w.p_instance, w.action, w.physics_param, w.state_mask,
                                 w.eef_mask, w.adj_thresh, w.topk, w.connect_tools_all, T)
    return preds


def time_cpu_baseline(params, Bc, steps, warmup, w=None):
    from adaptigraph_b200 import synthetic as syn
    torch.set_num_threads(os.cpu_count())
    if w is None:
        w = syn.make_workload(WORKLOAD["material"], WORKLOAD["n_p"], Bc, seed=1238)
    preds = None
    for _ in range(warmup):
        preds = cpu_rollout(w, params, WORKLOAD["pstep"], WORKLOAD["T"])
    t0 = time.perf_counter()
    for _ in range(steps):
        preds = cpu_rollout(w, params, WORKLOAD["pstep"], WORKLOAD["T"])
    dt = (time.perf_counter() - t0) / steps
    return Bc * WORKLOAD["n_p"] * WORKLOAD["T"] / dt, dt, preds, w


def cpu_model_name():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference(args, rank):
    if rank != 0:
        return
    from oracle import dynamics_oracle as orc
    params = orc.init_params(0)
    Bc = 1
    value, dt, _, _ = time_cpu_baseline(params, Bc, max(1, args.steps), max(0, args.warmup))
    sample = f"B={Bc} graph(s) of the cloth-2000 workload, T={WORKLOAD['T']} rollout per step (dense one-hot relations are O(B*E*N))"
    line = {
        "impl": "reference", "metric": "particle_steps_per_sec", "value": value, "unit": "particle-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_block(Bc),
        "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": os.cpu_count(), "kind": "port", "sample": sample,
                         "cpu": cpu_model_name()},
        "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def config_block(B):
    return {"workload": "cloth 2000 particles (+2 tool), 10-step rollout with per-step re-graph, pstep 3 (BASELINE configs[3])",
            "graphs_per_gpu": B, "n_p": WORKLOAD["n_p"], "n_tool": 2, "pstep": WORKLOAD["pstep"], "rollout_steps": WORKLOAD["T"],
            "adj_thresh": 0.75, "topk": 5, "connect_tools_all": True, "nf": F,
            "l2": "no flush: per-step working set (~2.6 GB of activations) exceeds the 126 MB L2"}


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.thread.join(timeout=6)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------ roofline accounting
def kernel_work(kind, B, N, n_p, E, K):
    """Algorithmic FLOPs and bytes of ONE launch of a kernel kind (derivations in DESIGN.md §5)."""
    rows = B * N
    if kind == "edge_encoder":          # 17->F->F->F relation encoder + the hoisted F->F relation part of the propagator
        return 2 * E * (17 * F + 3 * F * F), E * (8 + 2 * 64 + 4 * F)
    if kind == "node_encoder":          # 6->F->F->F + A_n, Qr, Qs products
        return 2 * rows * (6 * F + 5 * F * F), rows * (80 + 64 + 4 * 4 * F)
    if kind == "edge_aggregate":        # gather-add-relu-segmented-sum
        return 3 * E * F, E * (4 * F + 4) + rows * (3 * 4 * F + 8)
    if kind == "node_update":           # W_agg*agg + residual, then Qr, Qs
        return 2 * rows * 3 * F * F, rows * (6 * 4 * F)
    if kind == "node_update_head":      # W_agg*agg + residual, then the 3-layer motion head
        return 2 * rows * (3 * F * F + 3 * F), rows * (3 * 4 * F) + B * n_p * 36
    if kind == "graph_knn_rows":        # N^2 pair tests per graph, ~10 flops each
        return 10 * B * N * N, rows * (14 + 4 * 5 + 4)
    return 0, 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--graphs", type=int, default=WORKLOAD["B"], help="graphs per GPU (default: the BASELINE batch)")
    ap.add_argument("--total-graphs", type=int, default=0,
                    help="strong scaling: this many graphs in total, split evenly over the GPUs (overrides --graphs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--rollout-steps", type=int, default=WORKLOAD["T"], help="T (default: the BASELINE 10-step rollout)")
    args = ap.parse_args()

    WORKLOAD["T"] = args.rollout_steps
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl engine needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import __graft_entry__ as ge
    ge.build()
    import adaptigraph_b200 as agx
    from adaptigraph_b200 import ops, synthetic as syn

    strong = args.total_graphs > 0
    if strong:
        from adaptigraph_b200.shard import shard_slice
        sl = shard_slice(args.total_graphs, world, rank)
        args.graphs = sl.stop - sl.start
    B, n_p, K, T = args.graphs, WORKLOAD["n_p"], WORKLOAD["pstep"], WORKLOAD["T"]
    torch.manual_seed(0)
    model = agx.DynamicsPredictor(*syn.configs(WORKLOAD["material"], K), dev).to(dev).eval()
    params_cpu = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    # every rank gets its own shard of graphs (distinct seeds), graph 0.. of rank 0 = the CPU-baseline sample
    w_host = syn.make_workload(WORKLOAD["material"], n_p, B, seed=1238 + 1000 * rank)
    w = w_host.to(dev)
    N = w.N
    roll = lambda ww: model.rollout(ww.state, ww.attrs, ww.action, ww.p_instance, ww.physics_param, ww.state_mask,  # noqa: E731
                                    ww.eef_mask, ww.adj_thresh, ww.topk, ww.connect_tools_all, T, WORKLOAD["max_nR"], check=False)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput
    out = None
    for _ in range(max(args.warmup, 3)):
        out = roll(w)
    barrier()
    ops.profile_read()
    ops.profile_enable(True)
    launches0 = ops.launch_count()
    with ClockSampler(local_rank) as clocks:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(args.steps):
            out = roll(w)
        e1.record()
        barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = ops.launch_count() - launches0
    prof = ops.profile_read()
    ops.profile_enable(False)
    overflow = int(out["n_edges"].max().item()) > WORKLOAD["max_nR"]
    assert not overflow, "relation capacity exceeded in the bench workload"
    ms_per_step = ms_total / args.steps
    particle_steps = B * n_p * T
    job_particle_steps = (args.total_graphs if strong else world * B) * n_p * T      # all ranks together
    value = job_particle_steps / (ms_per_step * 1e-3)

    # ---- end to end from pinned host buffers (H2D of the step's inputs + D2H of the predictions each step)
    host = {k: v.pin_memory() for k, v in dict(state=w_host.state, attrs=w_host.attrs, action=w_host.action, p_instance=w_host.p_instance,
                                               physics_param=w_host.physics_param, state_mask=w_host.state_mask,
                                               eef_mask=w_host.eef_mask).items()}
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    out_host = torch.empty(B, T, n_p, 3, dtype=torch.float32).pin_memory()
    d2h = out_host.numel() * 4

    def e2e_step():
        d = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        o = model.rollout(d["state"], d["attrs"], d["action"], d["p_instance"], d["physics_param"], d["state_mask"], d["eef_mask"],
                          w_host.adj_thresh, w_host.topk, w_host.connect_tools_all, T, WORKLOAD["max_nR"], check=False)
        out_host.copy_(o["state_seqs"], non_blocking=True)

    for _ in range(2):
        e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    e2e_value = job_particle_steps / (e2e_ms * 1e-3)

    # ---- roofline of the dominant kernel (live CUDA-event timings of the timed region above)
    E = float(out["n_edges"].float().sum(1).mean().item())      # relations per model step over the batch
    pk = peaks()
    kernels = {}
    total_kernel_ms = sum(ms for ms, _ in prof.values())
    for kind, (ms, cnt) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        flops, byts = kernel_work(kind, B, N, n_p, E, K)
        per = ms / cnt
        kernels[kind] = {"launches": cnt, "avg_ms": per, "share": ms / total_kernel_ms,
                         "tflops": flops / per / 1e9 if flops else None, "gbs": byts / per / 1e6 if byts else None}
    top = next(iter(kernels))
    flops, byts = kernel_work(top, B, N, n_p, E, K)
    per = kernels[top]["avg_ms"]
    tensor_peak = pk["bf16_tflops_sustained"]          # 16-bit dense tensor peak measured inside a long step (cuBLAS bf16)
    intensity = flops / max(byts, 1)
    precision = os.environ.get("AGX_PRECISION", "tc")
    if intensity > tensor_peak * 1e12 / (pk["hbm_gbs"] * 1e9) / 3.0:
        # algorithmic FLOPs; the tc path executes 3 fp16 MMAs per product on 160-padded tiles (x3 x (160/150)^2)
        executed = flops / per / 1e9 * (3.0 * (160.0 / 150.0) ** 2 if precision == "tc" else 1.0)
        roof = {"kernel": top, "bound": "tensor", "achieved": flops / per / 1e9, "peak": tensor_peak, "unit": "TFLOP/s",
                "frac": flops / per / 1e9 / tensor_peak, "traffic": None, "executed_tensor_tflops": executed if precision == "tc" else None,
                "peak_source": pk["source"] + ": bf16_tflops_sustained; arithmetic = " +
                ("tcgen05 kind::f16, 3 split-fp16 MMAs per fp32-accurate product" if precision == "tc" else "fp32 FFMA tiles")}
    else:
        roof = {"kernel": top, "bound": "hbm", "achieved": byts / per / 1e6, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": byts / per / 1e6 / pk["hbm_gbs"], "traffic": None, "peak_source": pk["source"]}
    # DRAM traffic of that kernel per launch from the committed ncu --set full capture of this same command (profiles/ncu_traffic.json,
    # written by tools/ncu_summary.py); only valid for the default workload
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath) and B == WORKLOAD["B"] and n_p == WORKLOAD["n_p"]:
        tk = json.load(open(tpath))["kernels"].get(top)
        if tk:
            roof["traffic"] = tk["traffic_gb_per_launch"]
            roof["traffic_unit"] = "GB per launch (dram__bytes_read.sum + dram__bytes_write.sum; algorithmic: %.3f GB)" % (byts / 1e9)
    # whole-step view: stage-granular algorithmic bytes of SURVEY §8d over the measured step time
    Eg, = (E / B,)
    bytes_step = B * (N * (80 + 14 + 4 * F + K * 24 * F) + Eg * (16 + 4 * F + K * (4 * F + 8)) + n_p * (4 * F + 36) + 4)
    roof["step_hbm_frac"] = bytes_step * T / (ms_per_step * 1e-3) / 1e9 / pk["hbm_gbs"]
    roof["step_algorithmic_gb"] = bytes_step / 1e9

    line = {
        "metric": "particle_steps_per_sec", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": dict(config_block(B), arithmetic=os.environ.get("AGX_PRECISION", "tc")),
        "e2e": {"value": e2e_value, "unit": "particle-steps/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches, "roofline": roof, "kernels": kernels, "relations_per_graph": Eg,
        "clocks": clocks.summary(),
    }

    # ---- CPU baseline + parity on rank 0 at N=1 only
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        Bc = 1
        cpu_value, cpu_dt, cpu_preds, _ = time_cpu_baseline(params_cpu, Bc, steps=12, warmup=1, w=w_host.take(slice(0, Bc)))
        line["cpu_baseline"] = {"value": cpu_value, "unit": "particle-steps/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": f"B={Bc} graph of the same workload (graph 0), T={T} rollout, 12 timed passes of {cpu_dt:.1f} s each",
                                "cpu": cpu_model_name()}
        err = out["state_seqs"][:Bc].cpu() - cpu_preds
        line["parity"] = {"rollout_rmse_vs_cpu": float(err.pow(2).mean().sqrt()), "rollout_max_abs": float(err.abs().max()),
                          "sample": f"graph 0, all {T} steps, relations rebuilt on each side"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
