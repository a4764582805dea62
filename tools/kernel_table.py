"""Prints the headline and the per-kernel averages of a bench.py JSON line (argument: path)."""
import json
import sys

d = json.load(open(sys.argv[1]))
print(f'{d["value"] / 1e6:.2f} M {d["unit"]}   {d["ms_per_step"]:.3f} ms/step   rmse {d.get("parity", {}).get("rollout_rmse_vs_cpu")}')
for k, v in d["kernels"].items():
    print(f'  {k:18s} x{v["launches"]:4d}  {v["avg_ms"]:.4f} ms  share {v["share"]:.3f}')
