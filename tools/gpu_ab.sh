#!/bin/bash
# A/B of two builds on the same box: bench_train with variants/libagx_prev.so and with the in-tree library, twice each
T=$1
for i in 1 2; do
  AGX_LIB=$PWD/variants/libagx_prev.so timeout 300 python tests/bench/bench_train.py 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('prev', d['gpu_ms_per_step'], d['trainer_n_future3_ms_per_iter_cuda_graph'])" >> gpurun_out/${T}_ab.txt
  timeout 300 python tests/bench/bench_train.py 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('new ', d['gpu_ms_per_step'], d['trainer_n_future3_ms_per_iter_cuda_graph'])" >> gpurun_out/${T}_ab.txt
done
cat gpurun_out/${T}_ab.txt
