#!/bin/bash
# Session K: fp8-residual relation chain.  Usage: bash tools/gpu_r2k.sh TAG
T=${1:-r02k}; OUT=gpurun_out; mkdir -p $OUT
echo "== pytest parity"; timeout 1500 python -m pytest tests/test_parity_gpu.py tests/test_baseline_sizes_gpu.py -q -m gpu --tb=short -p no:cacheprovider > $OUT/${T}_pytest.log 2>&1; echo "rc=$?"; tail -25 $OUT/${T}_pytest.log | cut -c1-200
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
echo "== bench tc (fp8 residual)"; timeout 600 python bench.py > $OUT/${T}_bench_tc.json 2> $OUT/${T}_bench_tc.err; echo "rc=$?"; summ $OUT/${T}_bench_tc.json; tail -2 $OUT/${T}_bench_tc.err
echo "== bench tc (2 fp16 MMAs)"; AGX_LIB=adaptigraph_b200/libagx_mma2.so timeout 600 python bench.py > $OUT/${T}_bench_tc_mma2.json 2> $OUT/${T}_bench_tc_mma2.err; echo "rc=$?"; summ $OUT/${T}_bench_tc_mma2.json
echo "== bench cfg3"; timeout 600 python bench.py --workload cfg3 > $OUT/${T}_bench_cfg3.json 2> $OUT/${T}_bench_cfg3.err; echo "rc=$?"; summ $OUT/${T}_bench_cfg3.json
