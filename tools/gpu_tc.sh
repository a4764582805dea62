#!/bin/bash
# GPU session for the tensor-core path: per-stage diagnostics first (bounded waits trap instead of hanging),
# then the parity tests and a bench run with AGX_PRECISION=tc.  Stops early if the diagnostics fail.
TAG=${1:-r01b}
OUT=gpurun_out
mkdir -p $OUT
echo "== stage debug tc"; CUDA_LAUNCH_BLOCKING=1 AGX_DEBUG_PREC=1 timeout 120 python tests/bench/stage_debug.py > $OUT/${TAG}_stage_tc.log 2>&1; RC=$?; echo "rc=$RC"; tail -22 $OUT/${TAG}_stage_tc.log
if [ $RC -ne 0 ]; then
  echo "== compute-sanitizer (first failure)"; AGX_DEBUG_PREC=1 timeout 300 compute-sanitizer --tool memcheck python tests/bench/stage_debug.py forward_rope100_k1.npz 2>&1 | grep -v "^prec" | head -40
  exit 1
fi
echo "== pytest gpu"; timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -x > $OUT/${TAG}_pytest.log 2>&1; echo "rc=$?"; tail -25 $OUT/${TAG}_pytest.log
echo "== bench tc"; AGX_PRECISION=tc timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/${TAG}_bench_tc.json 2> $OUT/${TAG}_bench_tc.err; echo "rc=$?"; cat $OUT/${TAG}_bench_tc.json; tail -5 $OUT/${TAG}_bench_tc.err
echo "== bench train (cfg2)"; timeout 300 python tests/bench/bench_train.py > $OUT/${TAG}_bench_train.json 2> $OUT/${TAG}_bench_train.err; echo "rc=$?"; cat $OUT/${TAG}_bench_train.json; tail -3 $OUT/${TAG}_bench_train.err
