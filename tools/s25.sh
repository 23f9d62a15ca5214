#!/bin/bash
# two-GPU record of the final code: multi-process parity (all data planes), bench line at N = 2, CPU arm under torchrun
set -x
mkdir -p gpurun_out; O=gpurun_out
export NCCL_DEBUG=WARN
nvidia-smi --query-gpu=index,name --format=csv
timeout 900 python -u -m pytest tests/test_gpu_multiproc.py -m gpu -q -p no:cacheprovider > $O/pytest_gpu_multiproc.log 2>&1; tail -3 $O/pytest_gpu_multiproc.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29523 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; cat $O/bench_2gpu.json; tail -3 $O/bench_2gpu.err
timeout 300 $TR --master-port 29524 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/bench_ref_2gpu.json 2> $O/bench_ref_2gpu.err; cut -c1-300 $O/bench_ref_2gpu.json
