#!/bin/bash
# Session F: aggregate CTA-shape A/B, multi-CTA advance, 1-GPU sweep.  Usage: bash tools/gpu_r2f.sh TAG
T=${1:-r02f}; OUT=gpurun_out; mkdir -p $OUT
echo "== pytest gpu (parity + baseline sizes)"; timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_baseline_sizes_gpu.py -q -m gpu --tb=short -p no:cacheprovider > $OUT/${T}_pytest.log 2>&1; echo "rc=$?"; tail -5 $OUT/${T}_pytest.log | cut -c1-300
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
except Exception as e:
    print("no json", e); sys.exit(0)
print(sys.argv[1], "value %.1fM e2e %.1fM ms %.3f frac %.3f" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["ms_per_step"], d["roofline"]["step_hbm_frac"]))
print("   ", {k: round(v["avg_ms"], 4) for k, v in d["kernels"].items()})
PY
}
echo "== bench tc (4x4)"; timeout 600 python bench.py --no-cpu-baseline > $OUT/${T}_bench_tc.json 2> $OUT/${T}_bench_tc.err; echo "rc=$?"; summ $OUT/${T}_bench_tc.json; tail -2 $OUT/${T}_bench_tc.err
for V in 8x2 2x8 4x3; do
  echo "== bench tc aggregate $V"; AGX_LIB=adaptigraph_b200/libagx_a16_$V.so timeout 600 python bench.py --no-cpu-baseline > $OUT/${T}_bench_tc_$V.json 2> $OUT/${T}_bench_tc_$V.err; echo "rc=$?"; summ $OUT/${T}_bench_tc_$V.json
  echo "== bench cfg3 aggregate $V"; AGX_LIB=adaptigraph_b200/libagx_a16_$V.so timeout 600 python bench.py --workload cfg3 --no-cpu-baseline > $OUT/${T}_bench_cfg3_$V.json 2> $OUT/${T}_bench_cfg3_$V.err; echo "rc=$?"; summ $OUT/${T}_bench_cfg3_$V.json
done
echo "== bench 16 graphs"; timeout 600 python bench.py --graphs 16 --no-cpu-baseline > $OUT/${T}_bench_g16.json 2> $OUT/${T}_bench_g16.err; echo "rc=$?"; summ $OUT/${T}_bench_g16.json
echo "== sweep (1 GPU)"; timeout 900 python bench.py --sweep --steps 5 > $OUT/${T}_sweep.jsonl 2> $OUT/${T}_sweep.err; echo "rc=$?"; python - <<PY
import json
for l in open("$OUT/${T}_sweep.jsonl"):
    d = json.loads(l); c = d["config"]
    print("%-9s %5d x %4d  %.1fM  frac %.3f  E/graph %.0f" % (c["workload"].split()[0], c["n_p"], c["graphs_per_gpu"], d["value"] / 1e6, d["step_hbm_frac"], d["relations_per_graph"]))
PY
tail -2 $OUT/${T}_sweep.err
