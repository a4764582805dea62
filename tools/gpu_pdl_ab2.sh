#!/bin/bash
# Which part of programmatic dependent launch costs time?  AGX_PDL mask (1 = small kernels, 2 = chain / aggregate kernels) x
# {explicit early trigger, implicit trigger at CTA exit (variants/libagx_notrigger.so)} on the 128-graph bench rollout.
T=${1:-r02H}; OUT=gpurun_out; mkdir -p $OUT
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu-baseline --steps 10 --warmup 3 > $OUT/${T}_$name.json 2> $OUT/${T}_$name.err
  python - $OUT/${T}_$name.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); print(sys.argv[1], "value %.1fM ms %.3f" % (d["value"] / 1e6, d["ms_per_step"]))
except Exception as e:
    print(sys.argv[1], "no json", e)
PY
}
run trig_pdl0 AGX_PDL=0
run trig_pdl1 AGX_PDL=1
run trig_pdl2 AGX_PDL=2
run trig_pdl3 AGX_PDL=3
NT=$PWD/variants/libagx_notrigger.so
run notrig_pdl1 AGX_PDL=1 AGX_LIB=$NT
run notrig_pdl2 AGX_PDL=2 AGX_LIB=$NT
run notrig_pdl3 AGX_PDL=3 AGX_LIB=$NT
run trig_pdl0_again AGX_PDL=0

