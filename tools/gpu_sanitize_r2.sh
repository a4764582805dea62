#!/bin/bash
# compute-sanitizer over the final build: memcheck on the whole GPU suite, racecheck on the forward / rollout parity tests of the
# default arithmetic (C16 aggregate with fp32 row output and 7 slots, AGG32 update producer; long-row and ragged cases included).
# Usage: bash tools/gpu_sanitize_r2.sh TAG
T=${1:-r02san}; OUT=gpurun_out; mkdir -p $OUT
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -q -m gpu -p no:cacheprovider > $OUT/${T}_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "passed|failed|ERROR SUMMARY|Invalid" $OUT/${T}_memcheck.log | tail -6
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_parity_gpu.py -q -m gpu -p no:cacheprovider \
  -k "(long_rows or seeded or rollout_matches or without_relations) and tc and not tc3" > $OUT/${T}_racecheck.log 2>&1; echo "racecheck rc=$?"
grep -E "passed|failed|RACECHECK SUMMARY|hazard" $OUT/${T}_racecheck.log | tail -6
