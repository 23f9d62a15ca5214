#!/bin/bash
set -x
mkdir -p gpurun_out; O=gpurun_out
export NCCL_DEBUG=WARN
date +%T
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
D2D_PUSH=ce MGPU_SHAPES=small timeout 300 $TR --master-port 29521 tools/mgpu_check.py > $O/mgpu_ce_2.txt 2>&1; grep -E "^mgpu_check|FAIL|Error" $O/mgpu_ce_2.txt | head -5; tail -2 $O/mgpu_ce_2.txt
D2D_PUSH=ce MGPU_BACKEND=gloo D2D_TRANSPORT=boot MGPU_SHAPES=small CUDA_VISIBLE_DEVICES=0 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 tools/mgpu_check.py > $O/mgpu_ce_shared4.txt 2>&1; grep -E "^mgpu_check|FAIL|Error" $O/mgpu_ce_shared4.txt | head -5
date +%T
b() { name=$1; shift; timeout 300 env "$@" $TR --master-port 29530 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-cpu $BARGS > $O/bench2_$name.json 2> $O/bench2_$name.err; python - <<PY
import json
try:
    j=json.loads(open("$O/bench2_$name.json").read().strip().splitlines()[-1])
    r=j["roofline"]
    print("== $name ms/pair %.3f" % j["ms_per_step"], {k:v["ms_per_step"] for k,v in r["all_kernels"].items()}, r.get("exchanges"), r.get("flag_wait_ms_per_step"), j["forward_max_rel_err"])
except Exception as e:
    print("== $name FAILED", e); print(open("$O/bench2_$name.err").read()[-1500:])
PY
}
b ce_l2_k4 D2D_PUSH=ce D2D_COPY_LANES=2 D2D_CHUNKS=4
b ce_l1_k4 D2D_PUSH=ce D2D_COPY_LANES=1 D2D_CHUNKS=4
b ce_l2_k8 D2D_PUSH=ce D2D_COPY_LANES=2 D2D_CHUNKS=8
b ce_l2_k16 D2D_PUSH=ce D2D_COPY_LANES=2 D2D_CHUNKS=16
b ce_l1_k16 D2D_PUSH=ce D2D_COPY_LANES=1 D2D_CHUNKS=16
b ce_l2_k6e1 D2D_PUSH=ce D2D_COPY_LANES=2 D2D_CHUNKS=6 D2D_CHUNK_EDGE=1
date +%T
