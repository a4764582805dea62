#!/bin/bash
T=${1:-r02j}; OUT=gpurun_out; mkdir -p $OUT
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
except Exception as e:
    print("no json", e); sys.exit(0)
print(sys.argv[1], "value %.1fM ms %.3f" % (d["value"] / 1e6, d["ms_per_step"]))
print(" k:", {k: round(v["avg_ms"], 4) for k, v in d["kernels"].items()})
PY
}
for V in abl4 abl5; do
  echo "== bench tc $V"; AGX_LIB=adaptigraph_b200/libagx_$V.so timeout 600 python bench.py --no-cpu-baseline --steps 5 > $OUT/${T}_bench_tc_$V.json 2> $OUT/${T}_bench_tc_$V.err; echo "rc=$?"; summ $OUT/${T}_bench_tc_$V.json; tail -1 $OUT/${T}_bench_tc_$V.err
done
