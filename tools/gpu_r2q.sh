#!/bin/bash
T=$1
timeout 300 python tests/bench/grad_check.py granular 150 3 2 > gpurun_out/${T}_gradcheck.txt 2>&1
timeout 300 python tests/bench/bench_train.py > gpurun_out/${T}_train.json 2> gpurun_out/${T}_train.err
cat gpurun_out/${T}_train.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 300 --csv --log-file gpurun_out/${T}_train_launches.csv python tests/bench/bench_train.py > gpurun_out/${T}_train_ncu.log 2>&1
