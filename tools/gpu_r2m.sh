#!/bin/bash
# Session M: endpoint prefetch in the relation encoder, S0 for propagation step 0.  Usage: bash tools/gpu_r2m.sh TAG
T=${1:-r02m}; OUT=gpurun_out; mkdir -p $OUT
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > $OUT/${T}_pytest.log 2>&1; echo "rc=$?"; tail -8 $OUT/${T}_pytest.log | cut -c1-300
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
except Exception as e:
    print("no json", e); sys.exit(0)
print(sys.argv[1], "value %.1fM e2e %.1fM ms %.3f frac %.3f parity %s" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["ms_per_step"], d["roofline"]["step_hbm_frac"], d.get("parity")))
print(" k:", {k: round(v["avg_ms"], 4) for k, v in d["kernels"].items()})
PY
}
echo "== bench tc"; timeout 600 python bench.py > $OUT/${T}_bench_tc.json 2> $OUT/${T}_bench_tc.err; echo "rc=$?"; summ $OUT/${T}_bench_tc.json; tail -2 $OUT/${T}_bench_tc.err
echo "== bench tc3"; AGX_PRECISION=tc3 timeout 600 python bench.py --no-cpu-baseline > $OUT/${T}_bench_tc3.json 2> $OUT/${T}_bench_tc3.err; echo "rc=$?"; summ $OUT/${T}_bench_tc3.json
echo "== bench cfg3"; timeout 600 python bench.py --workload cfg3 --no-cpu-baseline > $OUT/${T}_bench_cfg3.json 2> $OUT/${T}_bench_cfg3.err; echo "rc=$?"; summ $OUT/${T}_bench_cfg3.json
echo "== bench 16 graphs"; timeout 600 python bench.py --graphs 16 --no-cpu-baseline > $OUT/${T}_bench_g16.json 2> $OUT/${T}_bench_g16.err; echo "rc=$?"; summ $OUT/${T}_bench_g16.json
