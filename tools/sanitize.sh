#!/bin/bash
# compute-sanitizer passes over the hand-written kernels (run under gpurun, 1 GPU).  The reference has no sanitizer
# configuration at all; this is the B200 path's own hygiene check.  Small shapes: memcheck and
# racecheck slow kernels down by 10-100x.  Logs: gpurun_out/sanitize_*.log
mkdir -p gpurun_out
T="tests/test_gpu_ops.py::test_attention tests/test_gpu_ops.py::test_gemm_epilogues tests/test_gpu_shots.py::test_forward_matches_reference_outputs tests/test_gpu_dedup.py::test_rowdot_argmax tests/test_gpu_preprocess.py::test_tensor_pipe_preprocess_agrees_with_simt_kernel_and_oracle tests/test_gpu_preprocess.py::test_resize_cubic_matches_cv2_goldens tests/test_gpu_preprocess.py::test_video_tube_from_nv12_surfaces"
for tool in ${SANITIZE_TOOLS:-memcheck racecheck}; do
  timeout ${SANITIZE_TIMEOUT:-600} compute-sanitizer --tool $tool --error-exitcode 9 --log-file gpurun_out/sanitize_$tool.log \
    python -m pytest $T -m gpu -x -q -k "${SANITIZE_K:-257-16-64 or epilogues or tiny7 or 300-300 or 360-640 or resize_cubic or video_tube}" > gpurun_out/sanitize_${tool}_pytest.log 2>&1
  echo "$tool: exit $?"; tail -3 gpurun_out/sanitize_$tool.log
done
