#!/bin/bash
TAG=${1:-r02r}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/${TAG}_build.log 2>&1
for d in 1024 1152 768; do for v in 0 1 2; do LN_D=$d CB_LN_VARIANT=$v python tools/ln_ab.py 2>/dev/null | tail -1; done; done | tee gpurun_out/${TAG}_ln_ab.jsonl
