#!/bin/bash
# eight-GPU record of the final code: multi-process parity at 8 ranks, the bench line exactly as the driver launches it
set -x
mkdir -p gpurun_out; O=gpurun_out
export NCCL_DEBUG=WARN
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
MGPU_SHAPES=small timeout 150 $TR8 --master-port 29521 tools/mgpu_check.py > $O/mgpu_8.txt 2>&1; grep -E "^mgpu_check|FAIL|Error" $O/mgpu_8.txt | head -5
timeout 240 $TR8 --master-port 29530 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_8gpu.json 2> $O/bench_8gpu.err; cat $O/bench_8gpu.json; tail -3 $O/bench_8gpu.err
