#!/bin/bash
# precision of the opt-in tensor-core training layers, one forward layer at a time
T=$1
for l in 1 2 4 5 6 7 8 9 10 11 12; do
  echo "=== layer $l" >> gpurun_out/${T}_gradcheck_layers.txt
  AGX_TRAIN_PRECISION=tc_one AGX_TRAIN_TC_LAYER=$l timeout 300 python tests/bench/grad_check.py granular 150 3 2 2>&1 | grep -E "state|relation_propagator.linear.weight|relation_encoder.model.0.weight|particle_encoder.model.0.weight" >> gpurun_out/${T}_gradcheck_layers.txt
done
