#!/bin/bash
# Bench-only call: the default line (all rows).  usage: gpu_call_bench.sh <tag>
TAG=${1:-r02_final3}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/${TAG}_build.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"; tail -c 300 gpurun_out/${TAG}_bench.err
