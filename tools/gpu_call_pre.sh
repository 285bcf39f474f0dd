#!/bin/bash
# Focused call: preprocess / tower tests (tensor-pipe kernel), value bench, ncu timing + full capture of the preprocess kernels.
TAG=${1:-r02e}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/${TAG}_build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_preprocess.py tests/test_gpu_tower.py -m gpu -x -q -s > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log; tail -4 gpurun_out/${TAG}_pytest.log
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-shots --no-cpu-baseline --no-gpu-library --no-secondary"
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-shots --no-cpu-baseline --no-gpu-library --no-secondary > gpurun_out/${TAG}_bench_value.json 2> gpurun_out/${TAG}_bench_value.err
echo "bench exit $?"
timeout 400 ncu --kernel-name regex:"preprocess|pack_patches" --metrics gpu__time_duration.sum --clock-control none -c 24 --csv --log-file gpurun_out/${TAG}_pre_launches.csv $B > /dev/null 2>&1
tail -8 gpurun_out/${TAG}_pre_launches.csv
timeout 400 ncu --kernel-name regex:"clip_preprocess_tc" --set full --import-source on --clock-control none -s 2 -c 1 -o gpurun_out/${TAG}_prof_pre_tc -f $B > /dev/null 2>&1
ls -la gpurun_out/${TAG}_prof_pre_tc.ncu-rep
if [ -n "$PRE_AB" ]; then
  CB_PRE_KERNEL=2 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-shots --no-cpu-baseline --no-gpu-library --no-secondary > gpurun_out/${TAG}_bench_value_v2.json 2>/dev/null
fi
