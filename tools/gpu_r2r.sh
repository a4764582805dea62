#!/bin/bash
T=$1
AGX_LIB=$PWD/variants/libagx_wgtl.so timeout 300 python tests/bench/grad_check.py granular 150 3 2 > gpurun_out/${T}_tl_small.txt 2>&1
AGX_LIB=$PWD/variants/libagx_wgtl.so timeout 300 python - > gpurun_out/${T}_tl.txt 2>&1 <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
sys.argv = ["bench_train.py"]
import runpy
try:
    runpy.run_path("tests/bench/bench_train.py", run_name="__main__")
except SystemExit:
    pass
PY
grep -c stage gpurun_out/${T}_tl.txt
