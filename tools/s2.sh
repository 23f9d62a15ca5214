#!/bin/bash
# Two-GPU session: parity of the three data planes, then the bench line with each of them / several chunk counts
set -x
mkdir -p gpurun_out; O=gpurun_out
export NCCL_DEBUG=WARN
nvidia-smi --query-gpu=index,name --format=csv
nvidia-smi topo -m | head -8
date +%T
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
D2D_P2P=1 timeout 300 $TR --master-port 29521 tools/mgpu_check.py > $O/mgpu_pipe_2.txt 2>&1; grep -E "^mgpu_check|FAIL|Error" $O/mgpu_pipe_2.txt | head -5
D2D_P2P=0 timeout 300 $TR --master-port 29522 tools/mgpu_check.py > $O/mgpu_nccl_2.txt 2>&1; grep -E "^mgpu_check|FAIL|Error" $O/mgpu_nccl_2.txt | head -5
D2D_FUSED=1 timeout 300 $TR --master-port 29523 tools/mgpu_check.py > $O/mgpu_fused_2.txt 2>&1; grep -E "^mgpu_check|FAIL|Error" $O/mgpu_fused_2.txt | head -5
date +%T
b() { name=$1; shift; timeout 300 env "$@" $TR --master-port 29530 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-cpu > $O/bench2_$name.json 2> $O/bench2_$name.err; python - <<PY
import json
try:
    j=json.loads(open("$O/bench2_$name.json").read().strip().splitlines()[-1])
    r=j["roofline"]
    print("== $name ms/pair %.3f" % j["ms_per_step"], {k:v["ms_per_step"] for k,v in r["all_kernels"].items()}, r.get("exchanges"), r.get("flag_wait_ms_per_step"))
except Exception as e:
    print("== $name FAILED", e); print(open("$O/bench2_$name.err").read()[-1500:])
PY
}
b k4 D2D_CHUNKS=4
b k8 D2D_CHUNKS=8
b k2 D2D_CHUNKS=2
b k1 D2D_CHUNKS=1
b k16 D2D_CHUNKS=16
b fused D2D_FUSED=1
b nccl4 D2D_P2P=0 D2D_CHUNKS=4
b nccl0 D2D_P2P=0 D2D_CHUNKS=0
date +%T
timeout 400 $TR --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench2_full.json 2> $O/bench2_full.err; cat $O/bench2_full.json; tail -3 $O/bench2_full.err
date +%T
