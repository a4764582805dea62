#!/bin/bash
T=$1
AGX_LIB=$PWD/variants/libagx_wgtl.so timeout 300 python tests/bench/grad_check.py rope 300 32 4 > gpurun_out/${T}_tl.txt 2>&1
timeout 300 python tests/bench/bench_train.py > gpurun_out/${T}_train.json 2> gpurun_out/${T}_train.err
cat gpurun_out/${T}_train.json
