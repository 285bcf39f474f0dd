#!/bin/bash
# Focused call: the new video-tube tests, sanitizer over the new kernel, then the full suite and a bench line with the iv2 row.
TAG=${1:-r02q}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/${TAG}_build.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -s -k "video_tube or internvideo2" > gpurun_out/${TAG}_iv2_pytest.log 2>&1
echo "iv2 pytest exit $?" >> gpurun_out/${TAG}_iv2_pytest.log; tail -15 gpurun_out/${TAG}_iv2_pytest.log
SANITIZE_TIMEOUT=300 SANITIZE_K="video_tube" bash tools/sanitize.sh
cp gpurun_out/sanitize_memcheck.log gpurun_out/${TAG}_sanitize_memcheck_video_tube.log 2>/dev/null
cp gpurun_out/sanitize_racecheck.log gpurun_out/${TAG}_sanitize_racecheck_video_tube.log 2>/dev/null
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log; tail -3 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-library > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"; tail -c 400 gpurun_out/${TAG}_bench.err
