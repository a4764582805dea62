#!/bin/bash
# Fused degrees + scan + fill kernel (profiles/r02L_experiment_fused_scan_fill.patch applied: scan_fill_rows_kernel) against the
# two-kernel path (AGX_GRAPH_FUSED_FILL=0): full GPU suite, then the bench at 128 and at 16 graphs.  Usage: bash tools/gpu_fill_ab.sh TAG
T=${1:-fill}; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -x > $OUT/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/${T}_pytest.log | cut -c1-300
summ() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
except Exception as e:
    print(sys.argv[2], "no json", e); sys.exit(0)
k = d["kernels"]
g = lambda n: k.get(n, {"avg_ms": 0.0})["avg_ms"]
print("%-14s value %.1fM e2e %.1fM launches/step %s | sort %.4f knn %.4f scan %.4f fill %.4f adv %.4f" % (sys.argv[2], d["value"] / 1e6, d["e2e"]["value"] / 1e6,
      d.get("launches_per_model_step"), g("graph_sort_cells"), g("graph_knn_rows"), g("graph_scan"), g("graph_fill_rows"), g("rollout_advance")))
PY
}
for rep in 1 2; do
  for v in 0 1; do
    AGX_GRAPH_FUSED_FILL=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench_f${v}_$rep.json 2> $OUT/${T}_bench_f${v}_$rep.err; summ $OUT/${T}_bench_f${v}_$rep.json fused$v
    AGX_GRAPH_FUSED_FILL=$v timeout 300 python bench.py --graphs 16 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/${T}_g16_f${v}_$rep.json 2> $OUT/${T}_g16_f${v}_$rep.err; summ $OUT/${T}_g16_f${v}_$rep.json g16-fused$v
  done
done
