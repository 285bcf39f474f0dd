#!/bin/bash
# Focused call: preprocess / tower tests (new tensor-pipe kernel), sanitizer on them, NVDEC session sweep.
TAG=${1:-r02d}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/${TAG}_build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_preprocess.py tests/test_gpu_tower.py -m gpu -x -q -s > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log; tail -4 gpurun_out/${TAG}_pytest.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/${TAG}_san_mem.log python -m pytest tests/test_gpu_preprocess.py -m gpu -x -q -k "tensor_pipe and 480 or resize_cubic_matches" > gpurun_out/${TAG}_san_mem_pytest.log 2>&1
echo "memcheck exit $?"; tail -2 gpurun_out/${TAG}_san_mem.log
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/${TAG}_san_race.log python -m pytest tests/test_gpu_preprocess.py -m gpu -x -q -k "tensor_pipe and 360" > gpurun_out/${TAG}_san_race_pytest.log 2>&1
echo "racecheck exit $?"; tail -2 gpurun_out/${TAG}_san_race.log
timeout 300 python tools/decode_sweep.py 8 14 20 28 40 > gpurun_out/${TAG}_decode_sweep.json 2> gpurun_out/${TAG}_decode_sweep.err
cat gpurun_out/${TAG}_decode_sweep.json
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-shots --no-cpu-baseline --no-gpu-library --no-secondary > gpurun_out/${TAG}_bench_value.json 2> gpurun_out/${TAG}_bench_value.err
echo "bench exit $?"
