#!/bin/bash
# Environment probe for the B200 box. Writes gpurun_out/probe.txt
mkdir -p gpurun_out
{
echo "== nvidia-smi"; nvidia-smi
echo "== nvidia-smi -q (encoder/decoder lines)"; nvidia-smi -q | grep -i -E 'decoder|encoder|jpeg|ofa|product name|cuda version' | head -40
echo "== libs"; ldconfig -p | grep -i -E 'nvcuvid|nvidia-encode|libcuda\.|nvjpeg|libnvidia-ml|avcodec' 
find / -name 'libnvcuvid*' -not -path '/proc/*' 2>/dev/null
find / -name 'libnvidia-encode*' -not -path '/proc/*' 2>/dev/null
find / \( -name 'nvcuvid.h' -o -name 'cuviddec.h' \) -not -path '/proc/*' 2>/dev/null
echo "== env"; env | grep -i -E 'nvidia|cuda' 
echo "== cpu"; nproc; lscpu | head -20; free -g | head -2
echo "== tools"; which ffmpeg ffprobe ncu nvcc; 
echo "== python"; python - <<'PY'
import torch, ctypes
print("torch", torch.__version__, torch.cuda.is_available(), torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
p = torch.cuda.get_device_properties(0); print(p)
for lib in ["libnvcuvid.so.1","libnvcuvid.so","libnvidia-encode.so.1"]:
    try:
        ctypes.CDLL(lib); print(lib,"LOADS")
    except OSError as e: print(lib,"FAIL",e)
import cv2; print(cv2.__version__)
print([l for l in cv2.getBuildInformation().splitlines() if any(k in l for k in ("FFMPEG","avcodec","NVCUVID","CUDA"))])
PY
} > gpurun_out/probe.txt 2>&1
tail -5 gpurun_out/probe.txt
