#!/bin/bash
# Quick GPU check: parity tests, then a short bench with the per-kernel table.  Usage: bash tools/gpu_quick.sh TAG [bench args]
T=${1:-quick}; shift
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${T}_pytest.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/${T}_bench.json"))
print(d["value"], d["ms_per_step"], d.get("parity"))
for k, v in d["kernels"].items():
    print(f"{k:20s} {v['avg_ms']:.4f} x{v['launches']}")
PY
tail -3 gpurun_out/${T}_bench.err
