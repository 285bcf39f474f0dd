#!/bin/bash
# Round-end evidence run (1 GPU): full GPU test suite, the default bench line + reference arm, sanitizer, ncu launch list + full captures.
TAG=${1:-r02_final}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt 2>&1
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/${TAG}_smoke.log
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log; tail -3 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "bench ref exit $?"
SANITIZE_TIMEOUT=300 bash tools/sanitize.sh > gpurun_out/${TAG}_sanitize.log 2>&1; tail -4 gpurun_out/${TAG}_sanitize.log
bash tools/profile.sh > gpurun_out/${TAG}_profile.log 2>&1; tail -8 gpurun_out/${TAG}_profile.log
