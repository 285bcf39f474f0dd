#!/bin/bash
# Run on the B200 box: GPU parity tests, one pytest process per file (a sticky CUDA error in one file must
# not mask the others), logs into gpurun_out/.
mkdir -p gpurun_out
rc=0
for f in ${GPU_TEST_FILES:-tests/test_gpu_preprocess.py tests/test_gpu_ops.py tests/test_gpu_tower.py tests/test_gpu_decode.py tests/test_gpu_stages.py tests/test_gpu_shots.py tests/test_gpu_dedup.py}; do
  [ -f "$f" ] || continue
  echo "=== $f"
  timeout 900 python -m pytest "$f" -m gpu -q --timeout 240 --timeout-method thread "$@" 2>&1 | tee gpurun_out/$(basename $f .py).log | tail -25
  [ ${PIPESTATUS[0]} -eq 0 ] || rc=1
done
exit $rc
