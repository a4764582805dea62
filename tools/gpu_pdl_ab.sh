#!/bin/bash
# A/B of programmatic dependent launch (AGX_PDL=0 plain launches vs the default) on the bench rollout at 128 and 16 graphs, cfg3,
# the training bench, and the GPU test suite with PDL on.  Usage: bash tools/gpu_pdl_ab.sh TAG
T=${1:-r02G}; OUT=gpurun_out; mkdir -p $OUT
line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], "value %.1fM e2e %.1fM ms %.3f frac %.3f" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["ms_per_step"], d["roofline"]["step_hbm_frac"]))
except Exception as e:
    print(sys.argv[1], "no json", e)
PY
}
echo "== pytest gpu (PDL on)"; timeout 1500 python -m pytest tests -q -m gpu --tb=short -x -p no:cacheprovider > $OUT/${T}_pytest.log 2>&1; echo "rc=$?"; tail -5 $OUT/${T}_pytest.log | cut -c1-300
for pdl in 0 1 0 1; do
  for g in 128 16; do
    AGX_PDL=$pdl timeout 600 python bench.py --graphs $g --no-cpu-baseline > $OUT/${T}_bench_g${g}_pdl${pdl}.json 2> $OUT/${T}_bench_g${g}_pdl${pdl}.err; line $OUT/${T}_bench_g${g}_pdl${pdl}.json
  done
done
for pdl in 0 1; do
  AGX_PDL=$pdl timeout 600 python bench.py --workload cfg3 --no-cpu-baseline > $OUT/${T}_bench_cfg3_pdl${pdl}.json 2> $OUT/${T}_bench_cfg3_pdl${pdl}.err; line $OUT/${T}_bench_cfg3_pdl${pdl}.json
  AGX_PDL=$pdl timeout 300 python tests/bench/bench_train.py > $OUT/${T}_train_pdl${pdl}.json 2> $OUT/${T}_train_pdl${pdl}.err; cut -c1-330 $OUT/${T}_train_pdl${pdl}.json
done
