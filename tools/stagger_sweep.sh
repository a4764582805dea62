for d in 0 5000 10000 20000 40000; do
  AGX_LIB=adaptigraph_b200/libagx_stagger.so AGX_TC_STAGGER=$d python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r01s_stag_$d.json 2> gpurun_out/r01s_stag_$d.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r01s_stag_$d.json"))
print($d, round(d["value"]/1e6,2), {k: round(v["avg_ms"],3) for k,v in d["kernels"].items() if k in ("edge_encoder","node_update","node_update_head","node_encoder","edge_aggregate")})
PY
done
