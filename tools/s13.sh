#!/bin/bash
set -x
mkdir -p gpurun_out; O=gpurun_out
export NCCL_DEBUG=WARN
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
{
timeout 200 $TR --master-port 29541 tools/bench_transposes_mgpu.py 1024 1 2 5
timeout 200 $TR --master-port 29542 tools/bench_transposes_mgpu.py 1024 2 1 5
timeout 200 $TR --master-port 29543 tools/bench_transposes_mgpu.py 512 1 2 10
D2D_P2P=0 timeout 200 $TR --master-port 29544 tools/bench_transposes_mgpu.py 1024 1 2 5
} 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$" | tee $O/transposes_2gpu.txt
