#!/bin/bash
# Run on the B200 box: GPU parity tests (+ optional extra args), logs into gpurun_out/.
mkdir -p gpurun_out
tools/cuvid_probe > gpurun_out/cuvid_probe.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 240 --timeout-method thread "$@" 2>&1 | tee gpurun_out/pytest_gpu.log | tail -40
