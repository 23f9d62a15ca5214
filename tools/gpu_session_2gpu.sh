#!/bin/bash
# Two-GPU sanity session: multi-process parity (peer-store path and NCCL path) and the bench line at N = 2.
set -x
mkdir -p gpurun_out
O=gpurun_out
export NCCL_DEBUG=WARN
nvidia-smi --query-gpu=index,name --format=csv
date +%T
D2D_P2P=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/mgpu_check.py > $O/mgpu_p2p_2.txt 2>&1; grep -E "mgpu_check|FAIL|failures|Error" $O/mgpu_p2p_2.txt | head
D2D_P2P=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 tools/mgpu_check.py > $O/mgpu_nccl_2.txt 2>&1; grep -E "mgpu_check|FAIL|failures|Error" $O/mgpu_nccl_2.txt | head
date +%T
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; cat $O/bench_2gpu.json; tail -3 $O/bench_2gpu.err
D2D_P2P=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > $O/bench_2gpu_nccl.json 2> $O/bench_2gpu_nccl.err; cat $O/bench_2gpu_nccl.json; tail -3 $O/bench_2gpu_nccl.err
date +%T
timeout 200 python -u -m pytest tests/test_gpu_fft_any.py -m gpu -q --maxfail=10 -p no:cacheprovider --timeout=90 > $O/pytest_gpu_any.log 2>&1; tail -3 $O/pytest_gpu_any.log
timeout 200 python -u -m pytest tests/test_gpu_multiproc.py -m gpu -q -p no:cacheprovider > $O/pytest_gpu_multiproc.log 2>&1; tail -3 $O/pytest_gpu_multiproc.log
date +%T
