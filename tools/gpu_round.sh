#!/bin/bash
# One GPU-box session: smoke, parity tests, per-stage diagnostic, sanitizer, bench, ncu launch list.
# Usage (from the repo root on the box): bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "rc=$?"; tail -3 $OUT/${TAG}_smoke.log
echo "== stage debug"; timeout 300 python tests/bench/stage_debug.py > $OUT/${TAG}_stage.log 2>&1; echo "rc=$?"; tail -20 $OUT/${TAG}_stage.log
echo "== pytest gpu"; timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1; echo "rc=$?"; tail -15 $OUT/${TAG}_pytest.log
echo "== sanitizer"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python tests/bench/stage_debug.py forward_rope100_k1.npz > $OUT/${TAG}_memcheck.log 2>&1; echo "rc=$?"; tail -5 $OUT/${TAG}_memcheck.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "rc=$?"; cat $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2>&1; echo "rc=$?"; cat $OUT/${TAG}_bench_ref.json
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --graphs 128 > $OUT/${TAG}_ncu_bench.log 2>&1; echo "rc=$?"
