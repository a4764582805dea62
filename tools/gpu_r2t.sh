#!/bin/bash
# Same-box A/B of the bench: previous build (variants/libagx_prev.so) against the in-tree build, after the parity tests.
# Usage: bash tools/gpu_r2t.sh TAG   (build variants/libagx_prev.so from the commit to compare with: see profiles/README.md)
T=${1:-r02T}; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_baseline_sizes_gpu.py -q -m gpu --tb=short -p no:cacheprovider -x > $OUT/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/${T}_pytest.log | cut -c1-300
summ() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
except Exception as e:
    print(sys.argv[2], "no json", e); sys.exit(0)
k = d["kernels"]
print("%-10s value %.1fM e2e %.1fM frac %.3f rmse %.2e | agg %.4f enc %.4f upd %.4f head %.4f nenc %.4f" % (sys.argv[2], d["value"] / 1e6, d["e2e"]["value"] / 1e6,
      d["roofline"]["step_hbm_frac"], (d.get("parity") or {}).get("rollout_rmse_vs_cpu", float("nan")), k["edge_aggregate"]["avg_ms"], k["edge_encoder"]["avg_ms"],
      k["node_update"]["avg_ms"], k["node_update_head"]["avg_ms"], k["node_encoder"]["avg_ms"]))
PY
}
for rep in 1 2; do
  for v in prev new; do
    case $v in prev) L=$PWD/variants/libagx_prev.so;; new) L=$PWD/adaptigraph_b200/libadaptigraph_b200.so;; b7c5) L=$PWD/variants/libagx_b7c5.so;; esac
    X=""; [ $rep = 2 ] && X="--no-cpu-baseline"
    AGX_LIB=$L timeout 300 python bench.py --steps 5 --warmup 3 $X > $OUT/${T}_bench_${v}_$rep.json 2> $OUT/${T}_bench_${v}_$rep.err; summ $OUT/${T}_bench_${v}_$rep.json $v
  done
done
for v in prev new; do
  case $v in prev) L=$PWD/variants/libagx_prev.so;; new) L=$PWD/adaptigraph_b200/libadaptigraph_b200.so;; b7c5) L=$PWD/variants/libagx_b7c5.so;; esac
  AGX_LIB=$L timeout 300 python bench.py --workload cfg3 --steps 5 --warmup 3 > $OUT/${T}_cfg3_${v}.json 2> $OUT/${T}_cfg3_${v}.err; summ $OUT/${T}_cfg3_${v}.json cfg3-$v
done
