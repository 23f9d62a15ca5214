#!/bin/bash
# Eight-GPU session: parity at 8 ranks, headline at 8 and 4 GPUs with the data-plane variants, the other BASELINE configs
set -x
mkdir -p gpurun_out; O=gpurun_out
export NCCL_DEBUG=WARN
date +%T
nvidia-smi --query-gpu=index,name --format=csv | head -3
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
MGPU_SHAPES=small timeout 200 $TR8 --master-port 29521 tools/mgpu_check.py > $O/mgpu_pipe_8.txt 2>&1; grep -E "^mgpu_check|FAIL|Error" $O/mgpu_pipe_8.txt | head -5
date +%T
b() { n=$1; name=$2; shift; shift; TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"; timeout 200 env "$@" $TR --master-port 29530 bench.py --gpus $n --steps 10 --warmup 3 --no-cpu $BARGS > $O/bench${n}_$name.json 2> $O/bench${n}_$name.err; python - <<PY
import json
try:
    j=json.loads(open("$O/bench${n}_$name.json").read().strip().splitlines()[-1])
    r=j["roofline"]
    print("== $n $name ms/pair %.3f" % j["ms_per_step"], {k:v["ms_per_step"] for k,v in r["all_kernels"].items()}, r.get("exchanges"), r.get("flag_wait_ms_per_step"), j["forward_max_rel_err"], (j.get("e2e") or {}).get("ms_per_step"), (j.get("e2e") or {}).get("pcie_spans"))
except Exception as e:
    print("== $n $name FAILED", e); print(open("$O/bench${n}_$name.err").read()[-1500:])
PY
}
b 8 k4 D2D_CHUNKS=4
BARGS="--no-e2e" b 8 k2 D2D_CHUNKS=2
BARGS="--no-e2e" b 8 k6 D2D_CHUNKS=6
BARGS="--no-e2e" b 8 k3e1 D2D_CHUNKS=3 D2D_CHUNK_EDGE=1.0
BARGS="--no-e2e" b 8 fused D2D_FUSED=1
date +%T
BARGS="--no-e2e --config 512x" b 8 512x X=1
BARGS="--config 2048f32 --steps 5 --e2e-steps 3" b 8 2048f32 X=1
date +%T
b 4 k4 D2D_CHUNKS=4
BARGS="--no-e2e" b 4 k6 D2D_CHUNKS=6
BARGS="--no-e2e" b 4 fused D2D_FUSED=1
date +%T
