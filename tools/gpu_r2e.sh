#!/bin/bash
# Round-2 GPU session E: full GPU suite, aggregate CTA-shape A/B, ring-search kNN, fused reward tail, MPC bench.  Usage: bash tools/gpu_r2e.sh TAG
T=${1:-r02e}; OUT=gpurun_out; mkdir -p $OUT
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > $OUT/${T}_pytest.log 2>&1; echo "rc=$?"; tail -30 $OUT/${T}_pytest.log | cut -c1-300
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
except Exception as e:
    print("no json", e); sys.exit(0)
print(sys.argv[1], "value %.1fM e2e %.1fM ms %.3f frac %.3f launches/step %s parity %s" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["ms_per_step"], d["roofline"]["step_hbm_frac"], d.get("launches_per_model_step"), d.get("parity")))
for k, v in d["kernels"].items():
    print(f"  {k:20s} {v['avg_ms']:.4f} x{v['launches']}")
PY
}
echo "== bench tc (4 receivers x 4 CTAs)"; timeout 600 python bench.py --no-cpu-baseline > $OUT/${T}_bench_tc.json 2> $OUT/${T}_bench_tc.err; echo "rc=$?"; summ $OUT/${T}_bench_tc.json; tail -2 $OUT/${T}_bench_tc.err
for V in 8x2 2x8; do
  echo "== bench tc aggregate $V"; AGX_LIB=adaptigraph_b200/libagx_a16_$V.so timeout 600 python bench.py --no-cpu-baseline > $OUT/${T}_bench_tc_$V.json 2> $OUT/${T}_bench_tc_$V.err; echo "rc=$?"; summ $OUT/${T}_bench_tc_$V.json
done
echo "== bench cfg3"; timeout 600 python bench.py --workload cfg3 --no-cpu-baseline > $OUT/${T}_bench_cfg3.json 2> $OUT/${T}_bench_cfg3.err; echo "rc=$?"; summ $OUT/${T}_bench_cfg3.json
echo "== bench 16 graphs"; timeout 600 python bench.py --graphs 16 --no-cpu-baseline > $OUT/${T}_bench_g16.json 2> $OUT/${T}_bench_g16.err; echo "rc=$?"; summ $OUT/${T}_bench_g16.json
echo "== bench mpc"; timeout 600 python tests/bench/bench_mpc.py > $OUT/${T}_mpc.json 2> $OUT/${T}_mpc.err; echo "rc=$?"; cat $OUT/${T}_mpc.json; tail -2 $OUT/${T}_mpc.err
echo "== graph bench"; timeout 300 python tools/bench_graph.py > $OUT/${T}_graph.txt 2>&1; tail -12 $OUT/${T}_graph.txt
