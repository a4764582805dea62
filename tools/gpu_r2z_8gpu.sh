#!/bin/bash
# Final build on one 8-GPU box: the 1-GPU point, weak scaling (128 graphs per GPU) and strong scaling (128 graphs in total) at 8 GPUs.
T=${1:-r02z8}; OUT=gpurun_out; mkdir -p $OUT
run() { # n tag extra-args
  local n=$1; shift; local tag=$1; shift
  if [ "$n" = "1" ]; then timeout 300 python bench.py --gpus 1 --no-cpu-baseline "$@" > $OUT/${T}_${tag}.json 2> $OUT/${T}_${tag}.err
  else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --no-cpu-baseline "$@" > $OUT/${T}_${tag}.json 2> $OUT/${T}_${tag}.err; fi
  echo "rc=$? $tag"; python - $OUT/${T}_${tag}.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print("   n_gpus %d %s value %.1fM e2e %.1fM ms %.3f graphs/gpu %s" % (d["n_gpus"], d["scaling"], d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["ms_per_step"], d["config"]["graphs_per_gpu"]))
except Exception as e:
    print("   no json", e)
PY
}
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $OUT/${T}_gpus.txt 2>&1
run 1 weak_1
run 8 weak_8
run 8 strong_8 --total-graphs 128
run 4 strong_4 --total-graphs 128
