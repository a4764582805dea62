#!/bin/bash
# Final round-2 session: smoke, full GPU suite, benches (default, tc3, cfg3, 16 graphs, reference arm), launch list + ncu --set full
# capture of one model step, training bench + its launch list.  Usage: bash tools/gpu_r2final.sh TAG
T=${1:-r02F}; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${T}_gpu.txt 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${T}_smoke.log 2>&1; echo "rc=$?"; tail -3 $OUT/${T}_smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > $OUT/${T}_pytest.log 2>&1; echo "rc=$?"; tail -5 $OUT/${T}_pytest.log | cut -c1-300
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
except Exception as e:
    print("no json", e); sys.exit(0)
print(sys.argv[1], "value %.1fM e2e %.1fM ms %.3f frac %.3f launches/step %s parity %s clocks %s" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["ms_per_step"], d["roofline"]["step_hbm_frac"], d.get("launches_per_model_step"), d.get("parity"), d.get("clocks")))
for k, v in d["kernels"].items():
    print(f"  {k:20s} {v['avg_ms']:.4f} x{v['launches']}")
PY
}
echo "== bench tc"; timeout 600 python bench.py > $OUT/${T}_bench_tc.json 2> $OUT/${T}_bench_tc.err; echo "rc=$?"; summ $OUT/${T}_bench_tc.json; tail -2 $OUT/${T}_bench_tc.err
echo "== bench tc3"; AGX_PRECISION=tc3 timeout 600 python bench.py --no-cpu-baseline > $OUT/${T}_bench_tc3.json 2> $OUT/${T}_bench_tc3.err; echo "rc=$?"; summ $OUT/${T}_bench_tc3.json
echo "== bench cfg3"; timeout 600 python bench.py --workload cfg3 > $OUT/${T}_bench_cfg3.json 2> $OUT/${T}_bench_cfg3.err; echo "rc=$?"; summ $OUT/${T}_bench_cfg3.json
echo "== bench 16 graphs"; timeout 600 python bench.py --graphs 16 --no-cpu-baseline > $OUT/${T}_bench_g16.json 2> $OUT/${T}_bench_g16.err; echo "rc=$?"; summ $OUT/${T}_bench_g16.json
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${T}_bench_ref.json 2> $OUT/${T}_bench_ref.err; echo "rc=$?"; cut -c1-300 $OUT/${T}_bench_ref.json
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${T}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --eager > $OUT/${T}_ncu_bench.log 2>&1; echo "rc=$?"
echo "== ncu full: one model step"; timeout 900 ncu --set full --clock-control none --import-source on -k "regex:tc_|edge_aggregate|knn_rows" -s 20 -c 9 -f -o $OUT/${T}_prof python bench.py --steps 1 --warmup 1 --no-cpu-baseline --eager > $OUT/${T}_ncu_full.log 2>&1; echo "rc=$?"; tail -2 $OUT/${T}_ncu_full.log | cut -c1-300
echo "== train"; timeout 300 python tests/bench/bench_train.py > $OUT/${T}_train.json 2> $OUT/${T}_train.err; echo "rc=$?"; cut -c1-400 $OUT/${T}_train.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv --log-file $OUT/${T}_train_launches.csv python tests/bench/bench_train.py > $OUT/${T}_train_ncu.log 2>&1; echo "rc=$?"
