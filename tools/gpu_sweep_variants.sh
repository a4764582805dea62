#!/bin/bash
# BASELINE configs[4] sweep (one GPU) for several library variants on the same box: bash tools/gpu_sweep_variants.sh TAG name=path ...
T=$1; shift; OUT=gpurun_out; mkdir -p $OUT
for rep in 1 2; do
  for nv in new=$PWD/adaptigraph_b200/libadaptigraph_b200.so "$@"; do
    n=${nv%%=*}; L=${nv#*=}; case $L in /*) ;; *) L=$PWD/$L;; esac
    AGX_LIB=$L timeout 300 python bench.py --sweep --steps 5 > $OUT/${T}_sweep_${n}_$rep.jsonl 2> $OUT/${T}_sweep_${n}_$rep.err
    python - $OUT/${T}_sweep_${n}_$rep.jsonl $n <<'PY'
import json, sys
v = {}
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); c = d["config"]; v.setdefault(c["workload"].split()[0], []).append(d["value"] / 1e6)
print("%-6s " % sys.argv[2] + "  ".join("%s %s" % (m, " ".join("%.1f" % x for x in xs)) for m, xs in v.items()))
PY
  done
done
