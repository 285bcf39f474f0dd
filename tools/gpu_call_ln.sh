#!/bin/bash
TAG=${1:-r02r}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/${TAG}_build.log 2>&1
python tools/ln_ab.py 2> gpurun_out/${TAG}_ln_ab.err | tee gpurun_out/${TAG}_ln_ab.jsonl
