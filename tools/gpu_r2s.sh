#!/bin/bash
T=$1
timeout 900 python -m pytest tests/test_train_gpu.py -q -m gpu > gpurun_out/${T}_pytest_train.txt 2>&1
tail -3 gpurun_out/${T}_pytest_train.txt
timeout 300 python tests/bench/bench_train.py > gpurun_out/${T}_train.json 2> gpurun_out/${T}_train.err
cat gpurun_out/${T}_train.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv --log-file gpurun_out/${T}_train_launches.csv python tests/bench/bench_train.py > gpurun_out/${T}_train_ncu.log 2>&1
