#!/bin/bash
# ncu captures of the top kernels (one GPU, short command).  Usage: bash tools/gpu_ncu.sh TAG "regex" skip count
TAG=${1:-r01c}; RE=${2:-tc_edge_encoder_kernel}; SKIP=${3:-1}; CNT=${4:-1}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$RE" -s $SKIP -c $CNT -f -o $OUT/${TAG}_prof \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu.log 2>&1
echo "rc=$?"; tail -5 $OUT/${TAG}_ncu.log; ls -la $OUT/${TAG}_prof.ncu-rep
