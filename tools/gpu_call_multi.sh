#!/bin/bash
# Multi-GPU bench line (one process per GPU over NCCL), launched exactly like the driver does.  usage: gpu_call_multi.sh <N> <tag>
N=${1:-2}; TAG=${2:-r02_n$N}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/${TAG}_build.log 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 \
  > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench N=$N exit $?"; tail -c 400 gpurun_out/${TAG}_bench.err; head -c 600 gpurun_out/${TAG}_bench.json
