#!/bin/bash
# Same-box comparison of library variants on the default bench (and cfg3): bash tools/gpu_variants.sh TAG name=path [name=path ...]
# ("new" = the in-tree build is always included).  Prints one line per run: throughput and the per-kernel milliseconds.
T=$1; shift; OUT=gpurun_out; mkdir -p $OUT
summ() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
except Exception as e:
    print(sys.argv[2], "no json", e); sys.exit(0)
k = d["kernels"]
print("%-12s value %.1fM e2e %.1fM frac %.3f | agg %.4f enc %.4f upd %.4f head %.4f knn %.4f" % (sys.argv[2], d["value"] / 1e6, d["e2e"]["value"] / 1e6,
      d["roofline"]["step_hbm_frac"], k["edge_aggregate"]["avg_ms"], k["edge_encoder"]["avg_ms"],
      k["node_update"]["avg_ms"], k["node_update_head"]["avg_ms"], k["graph_knn_rows"]["avg_ms"]))
PY
}
for rep in 1 2; do
  for nv in new=$PWD/adaptigraph_b200/libadaptigraph_b200.so "$@"; do
    n=${nv%%=*}; L=${nv#*=}; case $L in /*) ;; *) L=$PWD/$L;; esac
    AGX_LIB=$L timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/${T}_bench_${n}_$rep.json 2> $OUT/${T}_bench_${n}_$rep.err; summ $OUT/${T}_bench_${n}_$rep.json $n
  done
done
for nv in new=$PWD/adaptigraph_b200/libadaptigraph_b200.so "$@"; do
  n=${nv%%=*}; L=${nv#*=}; case $L in /*) ;; *) L=$PWD/$L;; esac
  AGX_LIB=$L timeout 300 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/${T}_cfg3_${n}.json 2> $OUT/${T}_cfg3_${n}.err; summ $OUT/${T}_cfg3_${n}.json cfg3-$n
done
