#!/bin/bash
T=$1
AGX_TRAIN_PRECISION=tc_bwd timeout 300 python tests/bench/bench_train.py > gpurun_out/${T}_train.json 2> gpurun_out/${T}_train.err
cat gpurun_out/${T}_train.json
AGX_TRAIN_PRECISION=tc_bwd timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv --log-file gpurun_out/${T}_train_launches.csv python tests/bench/bench_train.py > gpurun_out/${T}_train_ncu.log 2>&1
