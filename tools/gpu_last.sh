#!/bin/bash
# Last check of a build: smoke, the whole GPU suite, the default bench, granular and 16 graphs.  Usage: bash tools/gpu_last.sh TAG
T=${1:-last}; OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${T}_smoke.log
timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > $OUT/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/${T}_pytest.log | cut -c1-300
summ() { python - "$1" "$2" <<'PY'
import json, sys
d = json.load(open(sys.argv[1])); k = d["kernels"]
print("%-8s value %.1fM e2e %.1fM frac %.3f roofline %.3f rmse %s | agg %.4f enc %.4f upd %.4f head %.4f" % (sys.argv[2], d["value"] / 1e6, d["e2e"]["value"] / 1e6,
      d["roofline"]["step_hbm_frac"], d["roofline"]["frac"], (d.get("parity") or {}).get("rollout_rmse_vs_cpu"), k["edge_aggregate"]["avg_ms"], k["edge_encoder"]["avg_ms"],
      k["node_update"]["avg_ms"], k["node_update_head"]["avg_ms"]))
PY
}
timeout 400 python bench.py > $OUT/${T}_bench_tc.json 2> $OUT/${T}_bench_tc.err; echo "bench rc=$?"; summ $OUT/${T}_bench_tc.json cfg4
timeout 300 python bench.py --workload cfg3 > $OUT/${T}_bench_cfg3.json 2> $OUT/${T}_bench_cfg3.err; summ $OUT/${T}_bench_cfg3.json cfg3
timeout 300 python bench.py --graphs 16 --no-cpu-baseline > $OUT/${T}_bench_g16.json 2> $OUT/${T}_bench_g16.err; summ $OUT/${T}_bench_g16.json g16
