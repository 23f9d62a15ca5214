#!/bin/bash
# Two-GPU session: chunk-edge variants, other configs through bench.py
set -x
mkdir -p gpurun_out; O=gpurun_out
export NCCL_DEBUG=WARN
date +%T
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
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
b e05_k4 D2D_CHUNKS=4 D2D_CHUNK_EDGE=0.5
b e05_k5 D2D_CHUNKS=5 D2D_CHUNK_EDGE=0.5
b e05_k6 D2D_CHUNKS=6 D2D_CHUNK_EDGE=0.5
b e03_k5 D2D_CHUNKS=5 D2D_CHUNK_EDGE=0.3
b e1_k4 D2D_CHUNKS=4 D2D_CHUNK_EDGE=1.0
BARGS="--config 512x" b 512x X=1
BARGS="--config 2048f32" b 2048f32 X=1
date +%T
