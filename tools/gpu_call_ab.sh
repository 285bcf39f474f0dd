#!/bin/bash
# A/B of the preprocess kernel's wait policy: value bench per setting (no rebuild needed).  usage: gpu_call_ab.sh <tag>
TAG=${1:-r02i}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/${TAG}_build.log 2>&1
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-shots --no-cpu-baseline --no-gpu-library --no-secondary"
for ns in 0 32 200 2000; do
  CB_PRE_WAIT_NS=$ns timeout 200 $B > gpurun_out/${TAG}_wait_${ns}.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_wait_${ns}.json"))
print("wait_ns=${ns}", "pre ms/step", round(d["roofline_other"]["preprocess"]["ms_per_step"],3), "value", round(d["value"],1))
PY
done
