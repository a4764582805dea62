#!/bin/bash
# GPU session for the tensor-core path: per-stage diagnostics first (bounded waits trap instead of hanging),
# then the parity tests and a bench run with AGX_PRECISION=tc.
TAG=${1:-r01b}
OUT=gpurun_out
mkdir -p $OUT
echo "== stage debug tc"; AGX_DEBUG_PREC=1 timeout 120 python tools/stage_debug.py > $OUT/${TAG}_stage_tc.log 2>&1; echo "rc=$?"; tail -22 $OUT/${TAG}_stage_tc.log
echo "== pytest gpu"; timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -x > $OUT/${TAG}_pytest.log 2>&1; echo "rc=$?"; tail -25 $OUT/${TAG}_pytest.log
echo "== bench tc"; AGX_PRECISION=tc timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/${TAG}_bench_tc.json 2> $OUT/${TAG}_bench_tc.err; echo "rc=$?"; cat $OUT/${TAG}_bench_tc.json; tail -5 $OUT/${TAG}_bench_tc.err
