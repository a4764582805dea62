#!/bin/bash
# Round-2 first GPU session: parity of the three arithmetic modes, then A/B benches.  Usage: bash tools/gpu_r2a.sh TAG
T=${1:-r02a}; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${T}_gpu.txt 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${T}_smoke.log 2>&1; echo "rc=$?"; tail -3 $OUT/${T}_smoke.log
echo "== pytest parity"; timeout 1200 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > $OUT/${T}_pytest.log 2>&1; echo "rc=$?"; tail -40 $OUT/${T}_pytest.log
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
except Exception as e:
    print("no json", e); sys.exit(0)
print(sys.argv[1], "value %.1fM ms %.3f parity %s" % (d["value"] / 1e6, d["ms_per_step"], d.get("parity")))
for k, v in d["kernels"].items():
    print(f"  {k:20s} {v['avg_ms']:.4f} x{v['launches']}")
PY
}
for P in tc3 tc; do
  echo "== bench $P"; AGX_PRECISION=$P timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/${T}_bench_$P.json 2> $OUT/${T}_bench_$P.err; echo "rc=$?"; summ $OUT/${T}_bench_$P.json; tail -2 $OUT/${T}_bench_$P.err
done
echo "== bench tc, 3 aggregate CTAs per SM"; AGX_LIB=adaptigraph_b200/libagx_a16c3.so AGX_PRECISION=tc timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench_tc_a16c3.json 2> $OUT/${T}_bench_tc_a16c3.err; echo "rc=$?"; summ $OUT/${T}_bench_tc_a16c3.json
echo "== ncu launches (tc)"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/${T}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${T}_ncu_bench.log 2>&1; echo "rc=$?"
echo "== ncu full (tc): one model step"; timeout 900 ncu --set full --clock-control none --import-source on -k "regex:tc_|edge_aggregate|knn_rows" -s 26 -c 9 -f -o $OUT/${T}_prof python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${T}_ncu_full.log 2>&1; echo "rc=$?"; tail -3 $OUT/${T}_ncu_full.log
