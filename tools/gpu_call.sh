#!/bin/bash
# One gpurun call: GPU tests, the default bench line, sanitizer passes.  Everything lands in gpurun_out/<tag>_*.
TAG=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt 2>&1
nvidia-smi topo -m >> gpurun_out/${TAG}_gpus.txt 2>&1
lscpu | head -25 >> gpurun_out/${TAG}_gpus.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/${TAG}_build.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
if [ -z "$SKIP_BENCH" ]; then
  timeout 900 python bench.py --steps ${BENCH_STEPS:-10} --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
  echo "bench exit $?"; tail -c 600 gpurun_out/${TAG}_bench.err
  timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
  echo "bench ref exit $?"
fi
if [ -n "$SANITIZE" ]; then
  SANITIZE_TIMEOUT=400 bash tools/sanitize.sh
  cp gpurun_out/sanitize_memcheck.log gpurun_out/${TAG}_sanitize_memcheck.log 2>/dev/null
  cp gpurun_out/sanitize_racecheck.log gpurun_out/${TAG}_sanitize_racecheck.log 2>/dev/null
fi
