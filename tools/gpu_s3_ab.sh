#!/bin/bash
# Three-slot relation encoder (profiles/r02S3_experiment_three_slot_edge_encoder.patch applied: tc_edge_encoder_s3_kernel) against the
# two-slot chain (AGX_EDGE_S3=0), same library, same box.  Usage: bash tools/gpu_s3_ab.sh TAG
T=${1:-pipe}; OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_baseline_sizes_gpu.py -q -m gpu --tb=short -p no:cacheprovider -x > $OUT/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/${T}_pytest.log | cut -c1-400
summ() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
except Exception as e:
    print(sys.argv[2], "no json", e); sys.exit(0)
k = d["kernels"]
print("%-12s value %.1fM e2e %.1fM frac %.3f rmse %s | agg %.4f enc %.4f upd %.4f head %.4f" % (sys.argv[2], d["value"] / 1e6, d["e2e"]["value"] / 1e6,
      d["roofline"]["step_hbm_frac"], (d.get("parity") or {}).get("rollout_rmse_vs_cpu"), k["edge_aggregate"]["avg_ms"], k["edge_encoder"]["avg_ms"],
      k["node_update"]["avg_ms"], k["node_update_head"]["avg_ms"]))
PY
}
for rep in 1 2; do
  for v in 0 1; do
    X=""; [ $rep = 2 ] && X="--no-cpu-baseline"
    AGX_EDGE_S3=$v timeout 300 python bench.py --steps 5 --warmup 3 $X > $OUT/${T}_bench_s3_${v}_$rep.json 2> $OUT/${T}_bench_s3_${v}_$rep.err; summ $OUT/${T}_bench_s3_${v}_$rep.json s3_$v; tail -2 $OUT/${T}_bench_s3_${v}_$rep.err | cut -c1-300
  done
done
for v in 0 1; do
  AGX_EDGE_S3=$v timeout 300 python bench.py --workload cfg3 --steps 5 --warmup 3 > $OUT/${T}_cfg3_s3_${v}.json 2> $OUT/${T}_cfg3_s3_${v}.err; summ $OUT/${T}_cfg3_s3_${v}.json cfg3-s3_$v
done
