#!/bin/bash
set -x
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -u -m pytest tests/test_gpu_halo.py -m gpu -q --maxfail=5 -p no:cacheprovider --timeout=300 > $O/pytest_s12.log 2>&1; tail -8 $O/pytest_s12.log
MGPU_BACKEND=gloo D2D_TRANSPORT=boot MGPU_SHAPES=small CUDA_VISIBLE_DEVICES=0 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 tools/mgpu_check.py > $O/mgpu_s12.txt 2>&1; grep -E "^mgpu_check|FAIL|Error" $O/mgpu_s12.txt | head -5
