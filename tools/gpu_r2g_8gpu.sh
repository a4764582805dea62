#!/bin/bash
# 8-GPU box: weak and strong scaling of BASELINE configs[3] at 2 / 4 / 8 GPUs, the 1-GPU points on the same box, the configs[4] sweep at 8.
T=${1:-r02g}; OUT=gpurun_out; mkdir -p $OUT
run() { # n extra-args tag
  local n=$1; shift; local tag=$1; shift
  if [ "$n" = "1" ]; then timeout 600 python bench.py --gpus 1 --no-cpu-baseline "$@" > $OUT/${T}_${tag}.json 2> $OUT/${T}_${tag}.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --no-cpu-baseline "$@" > $OUT/${T}_${tag}.json 2> $OUT/${T}_${tag}.err; fi
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
for n in 1 2 4 8; do run $n weak_$n; done
for n in 2 4 8; do run $n strong_$n --total-graphs 128; done
echo "== sweep at 8 GPUs"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus 8 --sweep --steps 5 > $OUT/${T}_sweep_8gpu.jsonl 2> $OUT/${T}_sweep_8gpu.err; echo "rc=$?"; python - <<PY
import json
for l in open("$OUT/${T}_sweep_8gpu.jsonl"):
    if not l.startswith("{"): continue
    d = json.loads(l); c = d["config"]
    print("%-9s %5d x %4d/gpu  %.1fM" % (c["workload"].split()[0], c["n_p"], c["graphs_per_gpu"], d["value"] / 1e6))
PY
echo "== 2-GPU data-parallel training check"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29660 tools/check_train_dp.py > $OUT/${T}_train_dp.txt 2>&1; echo "rc=$?"; tail -3 $OUT/${T}_train_dp.txt
