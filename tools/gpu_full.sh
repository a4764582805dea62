#!/bin/bash
# Full GPU session: every GPU test, smoke, the bench (both arms), the ncu launch list and one --set full capture of the hot kernels.
# Usage (repo root on the box): bash tools/gpu_full.sh TAG
T=${1:-full}; O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${T}_gpu.txt 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; echo "rc=$?"; tail -2 $O/${T}_smoke.log
echo "== pytest gpu"; timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > $O/${T}_pytest.log 2>&1; echo "rc=$?"; tail -4 $O/${T}_pytest.log
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "rc=$?"; cut -c1-400 $O/${T}_bench.json; tail -2 $O/${T}_bench.err
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/${T}_bench_ref.json 2>$O/${T}_bench_ref.err; echo "rc=$?"; cut -c1-600 $O/${T}_bench_ref.json
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/${T}_ncu_bench.log 2>&1; echo "rc=$?"
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k "regex:edge_aggregate_split_kernel|knn_rows_kernel|tc_edge_encoder_kernel|tc_node_update_kernel|tc_node_encoder_kernel" -s 8 -c 8 -f -o $O/${T}_prof python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/${T}_ncu_full.log 2>&1; echo "rc=$?"; ls -la $O/${T}_prof.ncu-rep
