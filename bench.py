"""bench.py — particle-steps/s of the particle-graph dynamics hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl engine|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[3], the one the north-star target is quoted on and
that fits one GPU — cloth, 2000 object particles + 2 tool particles per graph, batch 128 graphs per GPU,
pstep 3, 10-step autoregressive rollout with the relations rebuilt every step.
One bench "step" = one such rollout over the batch; particle-steps = B * n_p * T.  With N GPUs every
rank rolls out its own 128 graphs (weak scaling, no data-path collective);
`--total-graphs 128` splits the BASELINE batch over the ranks instead (strong scaling).

Engine arm (default) prints one JSON line with
  value    device-resident throughput (inputs in HBM, CUDA events, max over ranks) of the rollout replayed as a CUDA graph
           (`agx.GraphedRollout`, the public API for repeated rollouts of one shape; `--eager` times the plain call)
  e2e      the same rollout through the public API from pinned HOST buffers: every step's H2D of its inputs and D2H of its
           predictions inside the timed region (double-buffered on two copy streams, as a caller with a stream of batches would)
  roofline the dominant kernel's achieved algorithmic B/s against MEASURED_PEAKS.json, from a SEPARATE profiled pass
           (per-kernel CUDA events are never enabled inside the timed regions)
  cpu_baseline the reference's own CPU path on the host cores, bounded sample
Reference arm (--impl reference) times that CPU path alone: the UNMODIFIED reference modules staged in oracle/_ref
(oracle/build_ref.py) when present (`kind: "reference"`), else the oracle port pinned to them by tests/golden (`kind: "port"`).

Other modes (not part of the driver contract): `--workload cfg3` (granular 1000 x 64, 5-step rollout), `--sweep` (BASELINE
configs[4]: rope / granular / cloth at 256 ... 8192 particles, 262144 particles per GPU, one JSON line per point).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    "cfg4": dict(material="cloth", n_p=2000, B=128, pstep=3, T=10, max_nR=16384,
                 name="cloth 2000 particles (+2 tool), 10-step rollout with per-step re-graph, pstep 3 (BASELINE configs[3])"),
    "cfg3": dict(material="granular", n_p=1000, B=64, pstep=3, T=5, max_nR=26000,
                 name="granular 1000 particles (+5 tool), 5-step rollout with per-step re-graph, pstep 3 (BASELINE configs[2])"),
}
WORKLOAD = dict(WORKLOADS["cfg4"])
F = 150
C16_ROW = 320      # bytes per relation of the per-relation term in the default ("tc") arithmetic: 16-bit block fixed point


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"], bf16_tflops_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------ CPU path
def _cpu_runner(params):
    """(kind, fn(workload, T) -> preds): the reference's own modules when oracle/_ref is staged, else the oracle port."""
    from adaptigraph_b200 import synthetic as syn
    try:
        from oracle import ref_runner
        if ref_runner.available():
            model = ref_runner.make_model(syn.configs(WORKLOAD["material"], WORKLOAD["pstep"]), params)
            return "reference", lambda w, T: ref_runner.rollout(model, w, T, WORKLOAD["max_nR"])[0]
    except Exception as e:  # noqa: BLE001  (a broken staging must not take the bench down: fall back to the pinned port, say so)
        sys.stderr.write(f"oracle/_ref unusable ({e!r}); timing the oracle port instead\n")
    from oracle import dynamics_oracle as orc

    def port(w, T):
        return orc.rollout_dense(params, WORKLOAD["pstep"], w.state, w.attrs, w.p_instance, w.action, w.physics_param, w.state_mask,
                                 w.eef_mask, w.adj_thresh, w.topk, w.connect_tools_all, T)[0]
    return "port", port


def time_cpu_baseline(params, Bc, steps, warmup, w=None, budget_s=None):
    """steps timed passes after `warmup` untimed ones; with budget_s the pass count is chosen from the (timed) warm-up pass so that
    the sample stays about budget_s seconds of CPU work."""
    from adaptigraph_b200 import synthetic as syn
    torch.set_num_threads(os.cpu_count())
    if w is None:
        w = syn.make_workload(WORKLOAD["material"], WORKLOAD["n_p"], Bc, seed=1238)
    kind, run = _cpu_runner(params)
    preds = None
    t0 = time.perf_counter()
    for _ in range(warmup):
        preds = run(w, WORKLOAD["T"])
    if budget_s and warmup:
        steps = max(2, min(steps, int(budget_s / max((time.perf_counter() - t0) / warmup, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(steps):
        preds = run(w, WORKLOAD["T"])
    dt = (time.perf_counter() - t0) / steps
    return Bc * WORKLOAD["n_p"] * WORKLOAD["T"] / dt, dt, preds, kind, steps


def cpu_model_name():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def golden_params():
    """The reference constructor's weights under torch.manual_seed(0) (tests/golden/weights_seed0.npz, written by the reference)."""
    import numpy as np
    with np.load(os.path.join(ROOT, "tests", "golden", "weights_seed0.npz")) as z:
        return {k: torch.from_numpy(z[k]) for k in z.files}


def run_reference(args, rank, world):
    if rank != 0:
        return
    params = golden_params()
    Bc = 1
    value, dt, _, kind, _ = time_cpu_baseline(params, Bc, max(1, args.steps), max(0, args.warmup))
    sample = (f"B={Bc} graph(s) of the workload, T={WORKLOAD['T']} rollout per step (the reference's dense one-hot relations are "
              f"O(B*E*N): the BASELINE batch does not fit; CPU throughput is batch-linear)")
    line = {
        "impl": "reference", "metric": "particle_steps_per_sec", "value": value, "unit": "particle-steps/s",
        # the contract's line (n_gpus = what was asked for); ONE CPU process runs whatever N is, so this arm does not scale with N and
        # a GPU-arm / reference-arm ratio at N > 1 only restates the N = 1 ratio times the GPU scaling
        "n_gpus": args.gpus, "cpu_processes": 1, "scales_with_gpus": False, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_block(Bc),
        "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": os.cpu_count(), "kind": kind, "sample": sample,
                         "cpu": cpu_model_name()},
        "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def config_block(B):
    from adaptigraph_b200 import synthetic as syn
    thr, topk, cta, n_tool = syn.MATERIALS[WORKLOAD["material"]]
    return {"workload": WORKLOAD["name"],
            "graphs_per_gpu": B, "n_p": WORKLOAD["n_p"], "n_tool": n_tool, "pstep": WORKLOAD["pstep"], "rollout_steps": WORKLOAD["T"],
            "adj_thresh": thr, "topk": topk, "connect_tools_all": cta, "nf": F,
            "l2": "no flush: per-step working set (> 1.5 GB of activations at the BASELINE batch) exceeds the 126 MB L2"}


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML in-process every 20 ms (nvidia-smi takes longer to start than a timed
    region lasts); falls back to polling nvidia-smi when NVML is unavailable."""

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
        except Exception:  # noqa: BLE001
            self.nvml = None

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v for v in vis.split(",") if v.strip() != ""]
            if index < len(ids) and ids[index].strip().isdigit():
                return int(ids[index])
        return index

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:  # noqa: BLE001
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        names = []
        for name, bit in (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4)):
            if r & bit:
                names.append(name)
        self.rows.append((sm, mx, names))

    def _sample_smi(self):
        import subprocess
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        if out:
            f = [x.strip() for x in out.split(",")]
            names = [n for n, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[2:6])
                     if v.lower().startswith("active")]
            self.rows.append((int(f[0]), int(f[1]), names))

    def _run(self):
        while not self.stop.is_set():
            try:
                self._sample_nvml() if self.nvml else self._sample_smi()
            except Exception:  # noqa: BLE001
                pass
            self.stop.wait(0.02 if self.nvml else 0.1)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.thread.join(timeout=6)

    def summary(self):
        sm = sorted(r[0] for r in self.rows)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_min_mhz": sm[0] if sm else None,
                "sm_max_mhz": max((r[1] for r in self.rows), default=None),
                "reasons": sorted({n for r in self.rows for n in r[2]}), "samples": len(self.rows),
                "how": "NVML every 20 ms over every timed region of this run (device-resident, end-to-end, profiled pass)"
                if self.nvml else "nvidia-smi polling"}


# ------------------------------------------------------------------------------------------ roofline accounting
def kernel_work(kind, B, N, n_p, E, K, c_row=4 * F):
    """Algorithmic FLOPs and bytes of ONE launch of a kernel kind (derivations in DESIGN.md §5).  c_row = bytes per relation of
    the stored per-relation term: 4F in fp32 (AGX_PRECISION=fp32 / tc3), 320 as 16-bit block fixed point (tc)."""
    rows = B * N
    if kind == "edge_encoder":          # 17->F->F->F relation encoder + the hoisted F->F relation part of the propagator
        return 2 * E * (17 * F + 3 * F * F), E * (8 + 2 * 64 + c_row)
    if kind == "node_encoder":          # 6->F->F->F + A_n, Qr, Qs products
        return 2 * rows * (6 * F + 5 * F * F), rows * (80 + 64 + 4 * 4 * F)
    if kind == "edge_aggregate":        # gather-add-relu-segmented-sum
        return 3 * E * F, E * (c_row + 4) + rows * (3 * 4 * F + 8)
    if kind == "node_update":           # W_agg*agg + residual, then Qr, Qs
        return 2 * rows * 3 * F * F, rows * (6 * 4 * F)
    if kind == "node_update_head":      # W_agg*agg + residual, then the 3-layer motion head
        return 2 * rows * (3 * F * F + 3 * F), rows * (3 * 4 * F) + B * n_p * 36
    if kind == "graph_knn_rows":        # candidate scan of the 9 neighbouring cells, ~10 flops per pair
        return 10 * B * N * N, rows * (14 + 4 * 5 + 4)
    return 0, 0


def step_algorithmic_bytes(B, N, n_p, Eg, K):
    """stage-granular fp32 bytes of one model step over the batch."""
    return B * (N * (80 + 14 + 4 * F + K * 24 * F) + Eg * (16 + 4 * F + K * (4 * F + 8)) + n_p * (4 * F + 36) + 4)


# ------------------------------------------------------------------------------------------ engine arm
def bench_engine(args, rank, world, local_rank, sweep_point=None):
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist
    import adaptigraph_b200 as agx
    from adaptigraph_b200 import ops, synthetic as syn

    strong = args.total_graphs > 0
    graphs = args.graphs
    if sweep_point is not None:
        WORKLOAD.update(material=sweep_point[0], n_p=sweep_point[1], B=262144 // sweep_point[1], pstep=3, T=10,
                        max_nR=sweep_point[1] * 30,
                        name=f"{sweep_point[0]} {sweep_point[1]} particles, 10-step rollout with per-step re-graph (BASELINE configs[4] sweep)")
        graphs = WORKLOAD["B"]
    elif strong:
        from adaptigraph_b200.shard import shard_slice
        sl = shard_slice(args.total_graphs, world, rank)
        graphs = sl.stop - sl.start
    B, n_p, K, T = graphs, WORKLOAD["n_p"], WORKLOAD["pstep"], WORKLOAD["T"]
    torch.manual_seed(0)
    model = agx.DynamicsPredictor(*syn.configs(WORKLOAD["material"], K), dev).to(dev).eval()
    model.load_state_dict(golden_params())          # = the reference constructor under manual_seed(0); both arms use these weights
    # every rank gets its own shard of graphs (distinct seeds), graph 0.. of rank 0 = the CPU-baseline sample
    w_host = syn.make_workload(WORKLOAD["material"], n_p, B, seed=1238 + 1000 * rank)
    w = w_host.to(dev)
    N = w.N
    roll_args = lambda ww: (ww.state, ww.attrs, ww.action, ww.p_instance, ww.physics_param, ww.state_mask, ww.eef_mask,  # noqa: E731
                            ww.adj_thresh, ww.topk, ww.connect_tools_all, T, WORKLOAD["max_nR"])
    eager = lambda ww: model.rollout(*roll_args(ww), check=False)  # noqa: E731

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

    eager(w)                                                           # first call: packs the weights
    launches0 = ops.launch_count()
    out = eager(w)                                                     # one eager rollout: kernels per rollout, overflow check
    torch.cuda.synchronize()
    launches_per_rollout = ops.launch_count() - launches0
    assert int(out["n_edges"].max().item()) <= WORKLOAD["max_nR"] and not int(out["status"].item()) & 1, \
        "relation capacity exceeded in the bench workload"
    graphed = None if args.eager else agx.GraphedRollout(model, *roll_args(w))
    step = (lambda: eager(w)) if graphed is None else (lambda: graphed())

    with ClockSampler(local_rank) as clocks:
        # ---- device-resident throughput
        for _ in range(max(args.warmup, 3)):
            out = step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(args.steps):
            out = step()
        e1.record()
        barrier()
        ms_total = max_over_ranks(e0.elapsed_time(e1))
        ms_per_step = ms_total / args.steps
        job_particle_steps = (args.total_graphs if strong else world * B) * n_p * T      # all ranks together
        value = job_particle_steps / (ms_per_step * 1e-3)
        result_state = out["state_seqs"].clone()

        # ---- end to end from pinned host buffers: every step uploads its inputs and reads its predictions back.  Steps are
        # pipelined the way a caller with a stream of host batches would drive the public API: two staging buffers each way, the
        # H2D of step i + 1 and the D2H of step i - 1 on their own streams under the rollout of step i.
        host = {k: v.pin_memory() for k, v in dict(state=w_host.state, attrs=w_host.attrs, action=w_host.action, p_instance=w_host.p_instance,
                                                   physics_param=w_host.physics_param, state_mask=w_host.state_mask,
                                                   eef_mask=w_host.eef_mask).items()}
        h2d = sum(v.numel() * v.element_size() for v in host.values())
        d2h = B * T * n_p * 3 * 4
        main = torch.cuda.current_stream()
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        dev_in = [{k: torch.empty_like(v, device=dev) for k, v in host.items()} for _ in range(2)]
        dev_out = [torch.empty(B, T, n_p, 3, dtype=torch.float32, device=dev) for _ in range(2)]
        host_out = [torch.empty(B, T, n_p, 3, dtype=torch.float32).pin_memory() for _ in range(2)]
        ev = lambda: [torch.cuda.Event() for _ in range(2)]  # noqa: E731
        in_ready, in_free, out_ready, out_free = ev(), ev(), ev(), ev()

        def e2e_step(i):
            b = i & 1
            with torch.cuda.stream(s_in):
                s_in.wait_event(in_free[b])                           # staging buffer b was consumed by step i - 2
                for k, v in host.items():
                    dev_in[b][k].copy_(v, non_blocking=True)          # H2D
                in_ready[b].record(s_in)
            main.wait_event(in_ready[b])
            if graphed is not None:
                o = graphed(**dev_in[b])                              # into the captured buffers, replay
            else:
                d = dev_in[b]
                o = model.rollout(d["state"], d["attrs"], d["action"], d["p_instance"], d["physics_param"], d["state_mask"], d["eef_mask"],
                                  w_host.adj_thresh, w_host.topk, w_host.connect_tools_all, T, WORKLOAD["max_nR"], check=False)
            in_free[b].record(main)
            main.wait_event(out_free[b])                              # dev_out[b] has left for the host (step i - 2)
            dev_out[b].copy_(o["state_seqs"])
            out_ready[b].record(main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(out_ready[b])
                host_out[b].copy_(dev_out[b], non_blocking=True)      # D2H
                out_free[b].record(s_out)

        for i in range(2):
            e2e_step(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            e2e_step(i)
        main.wait_stream(s_in)
        main.wait_stream(s_out)                                       # the last step's predictions are on the host
        e1.record()
        barrier()
        e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
        e2e_value = job_particle_steps / (e2e_ms * 1e-3)
        assert torch.equal(host_out[(args.steps - 1) & 1].to(dev), result_state), "the end-to-end pass must reproduce the device-resident result"

        # ---- separate profiled pass (eager, per-kernel CUDA events on the launch stream): never inside a timed region above
        ops.profile_read()
        ops.profile_enable(True)
        for _ in range(max(2, min(args.steps, 5))):
            eager(w)
        torch.cuda.synchronize()
        prof = ops.profile_read()
        ops.profile_enable(False)

    # ---- roofline of the dominant kernel
    precision = os.environ.get("AGX_PRECISION", "tc")
    c_row = C16_ROW if precision == "tc" else 4 * F
    E = float(out["n_edges"].float().sum(1).mean().item())      # relations per model step over the batch
    pk = peaks()
    kernels = {}
    total_kernel_ms = sum(ms for ms, _ in prof.values())
    for kind, (ms, cnt) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        flops, byts = kernel_work(kind, B, N, n_p, E, K, c_row)
        per = ms / cnt
        kernels[kind] = {"launches": cnt, "avg_ms": per, "share": ms / total_kernel_ms,
                         "tflops": flops / per / 1e9 if flops else None, "gbs": byts / per / 1e6 if byts else None}
    top = next(iter(kernels))
    flops, byts = kernel_work(top, B, N, n_p, E, K, c_row)
    per = kernels[top]["avg_ms"]
    tensor_peak = pk["bf16_tflops_sustained"]          # 16-bit dense tensor peak measured inside a long step (cuBLAS bf16)
    intensity = flops / max(byts, 1)
    if intensity > tensor_peak * 1e12 / (pk["hbm_gbs"] * 1e9) / 3.0:
        mmas = {"tc": 2.0, "tc3": 3.0}.get(precision, 1.0) if top == "edge_encoder" else (3.0 if precision in ("tc", "tc3") else 1.0)
        roof = {"kernel": top, "bound": "tensor", "achieved": flops / per / 1e9, "peak": tensor_peak, "unit": "TFLOP/s",
                "frac": flops / per / 1e9 / tensor_peak, "traffic": None,
                "executed_tensor_tflops": flops / per / 1e9 * mmas * (160.0 / 150.0) ** 2 if precision != "fp32" else None,
                "peak_source": pk["source"] + ": bf16_tflops_sustained; arithmetic = " +
                (f"tcgen05 kind::f16, {mmas:.0f} split-fp16 MMAs per product on 160-padded tiles" if precision != "fp32" else "fp32 FFMA tiles")}
    else:
        roof = {"kernel": top, "bound": "hbm", "achieved": byts / per / 1e6, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": byts / per / 1e6 / pk["hbm_gbs"], "traffic": None, "peak_source": pk["source"],
                "algorithmic_bytes": "bytes of the kernel's inputs read once + outputs written once in the formats it runs on"
                                     + (" (per-relation term as 320-byte 16-bit block rows)" if c_row == C16_ROW else "")}
    # DRAM traffic of that kernel per launch from the committed ncu --set full capture of this same command (profiles/ncu_traffic.json,
    # written by tools/ncu_summary.py); only valid for the default workload and arithmetic
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath) and B == WORKLOADS["cfg4"]["B"] and n_p == WORKLOADS["cfg4"]["n_p"] and precision == "tc":
        tj = json.load(open(tpath))
        tk = tj["kernels"].get(top)
        if tk and tj.get("arithmetic", "tc3") == precision:
            roof["traffic"] = tk["traffic_gb_per_launch"]
            roof["traffic_unit"] = "GB per launch (dram__bytes_read.sum + dram__bytes_write.sum; algorithmic: %.3f GB)" % (byts / 1e9)
    # whole-step view: stage-granular fp32 algorithmic bytes of SURVEY §8d over the measured step time
    Eg = E / B
    bytes_step = step_algorithmic_bytes(B, N, n_p, Eg, K)
    roof["step_hbm_frac"] = bytes_step * T / (ms_per_step * 1e-3) / 1e9 / pk["hbm_gbs"]
    roof["step_algorithmic_gb"] = bytes_step / 1e9

    line = {
        "metric": "particle_steps_per_sec", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(config_block(B), arithmetic=precision, launch="eager" if graphed is None else "cuda graph replay (GraphedRollout)"),
        "e2e": {"value": e2e_value, "unit": "particle-steps/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "how": "pinned host inputs uploaded and predictions read back every step; copies double-buffered on two side streams"},
        "gpu_launches": launches_per_rollout * args.steps, "launches_per_model_step": launches_per_rollout / T,
        "roofline": roof, "kernels": kernels, "relations_per_graph": Eg, "clocks": clocks.summary(),
    }

    # ---- CPU baseline + parity on rank 0 at N=1 only
    if rank == 0 and world == 1 and not args.no_cpu_baseline and sweep_point is None:
        Bc = 1
        cpu_value, cpu_dt, cpu_preds, kind, cpu_steps = time_cpu_baseline(golden_params(), Bc, steps=24, warmup=1, w=w_host.take(slice(0, Bc)),
                                                                          budget_s=20.0)
        line["cpu_baseline"] = {"value": cpu_value, "unit": "particle-steps/s", "cores": os.cpu_count(), "kind": kind,
                                "sample": f"B={Bc} graph of the same workload (graph 0), T={T} rollout, {cpu_steps} timed passes of {cpu_dt:.1f} s each",
                                "cpu": cpu_model_name()}
        err = result_state[:Bc].cpu() - cpu_preds
        line["parity"] = {"rollout_rmse_vs_cpu": float(err.pow(2).mean().sqrt()), "rollout_max_abs": float(err.abs().max()),
                          "sample": f"graph 0, all {T} steps, relations rebuilt on each side", "against": kind}
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--graphs", type=int, default=0, help="graphs per GPU (default: the workload's BASELINE batch)")
    ap.add_argument("--total-graphs", type=int, default=0,
                    help="strong scaling: this many graphs in total, split evenly over the GPUs (overrides --graphs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true", help="time the plain model.rollout call instead of the CUDA-graph replay")
    ap.add_argument("--rollout-steps", type=int, default=0, help="T (default: the workload's BASELINE rollout length)")
    ap.add_argument("--sweep", action="store_true", help="BASELINE configs[4]: 3 materials x 6 sizes, one JSON line per point")
    args = ap.parse_args()

    WORKLOAD.update(WORKLOADS[args.workload])
    if args.rollout_steps:
        WORKLOAD["T"] = args.rollout_steps
    if not args.graphs:
        args.graphs = WORKLOAD["B"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl engine needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import __graft_entry__ as ge
    ge.build()

    if args.sweep:
        for material in ("rope", "granular", "cloth"):
            for n_p in (256, 512, 1024, 2048, 4096, 8192):
                line = bench_engine(args, rank, world, local_rank, sweep_point=(material, n_p))
                if rank == 0:
                    keep = ("metric", "value", "unit", "n_gpus", "ms_per_step", "scaling", "config", "e2e", "relations_per_graph")
                    print(json.dumps({k: line[k] for k in keep} | {"step_hbm_frac": line["roofline"]["step_hbm_frac"],
                                                                   "kernels": {k: round(v["avg_ms"], 4) for k, v in line["kernels"].items()}}), flush=True)
    else:
        line = bench_engine(args, rank, world, local_rank)
        if rank == 0:
            print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
