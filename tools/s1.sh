#!/bin/bash
set -x
mkdir -p gpurun_out; O=gpurun_out
date +%T
nvidia-smi --query-gpu=index,name --format=csv
MGPU_BACKEND=gloo D2D_TRANSPORT=boot MGPU_SHAPES=small CUDA_VISIBLE_DEVICES=0 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/mgpu_check.py > $O/mgpu_shared_2.txt 2>&1; echo "rc=$?"; grep -E "mgpu_check|FAIL|failures|Error|error" $O/mgpu_shared_2.txt | head -20; tail -5 $O/mgpu_shared_2.txt
date +%T
timeout 600 python -u -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider --timeout=300 > $O/pytest_gpu.log 2>&1; tail -15 $O/pytest_gpu.log
date +%T
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -6 $O/smoke.log
date +%T
timeout 400 python -u bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err
cat $O/bench.json; tail -5 $O/bench.err
date +%T
