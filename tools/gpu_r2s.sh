#!/bin/bash
T=$1
timeout 900 python -m pytest tests/test_train_gpu.py -q -m gpu > gpurun_out/${T}_pytest_train.txt 2>&1
tail -30 gpurun_out/${T}_pytest_train.txt
