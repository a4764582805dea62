#!/bin/bash
# Session N: small-graph sort CTAs, pipelined end-to-end bench leg.  Usage: bash tools/gpu_r2n.sh TAG
T=${1:-r02n}; OUT=gpurun_out; mkdir -p $OUT
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > $OUT/${T}_pytest.log 2>&1; echo "rc=$?"; tail -8 $OUT/${T}_pytest.log | cut -c1-300
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
except Exception as e:
    print("no json", e); sys.exit(0)
print(sys.argv[1], "value %.1fM e2e %.1fM ms %.3f e2e_ms %.3f frac %.3f parity %s clocks %s" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["step_hbm_frac"], d.get("parity"), d.get("clocks")))
print(" k:", {k: round(v["avg_ms"], 4) for k, v in d["kernels"].items()})
PY
}
echo "== bench tc"; timeout 600 python bench.py > $OUT/${T}_bench_tc.json 2> $OUT/${T}_bench_tc.err; echo "rc=$?"; summ $OUT/${T}_bench_tc.json; tail -2 $OUT/${T}_bench_tc.err
echo "== bench mpc"; timeout 600 python tests/bench/bench_mpc.py > $OUT/${T}_mpc.json 2> $OUT/${T}_mpc.err; echo "rc=$?"; cat $OUT/${T}_mpc.json; tail -2 $OUT/${T}_mpc.err
echo "== graph bench"; timeout 300 python tools/bench_graph.py > $OUT/${T}_graph.txt 2>&1; tail -8 $OUT/${T}_graph.txt
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${T}_bench_ref.json 2> $OUT/${T}_bench_ref.err; echo "rc=$?"; cut -c1-300 $OUT/${T}_bench_ref.json
