#!/bin/bash
# NVLink evidence for the kernels that talk to a peer: rank 0 of a 2-rank bench run under ncu (few metrics, one pass),
# rank 1 plain.  The flags between the ranks are stream memory operations, not kernels, so the replay of rank 0's kernels
# only makes rank 1 wait.  usage: tools/ncu_rank0_2gpu.sh <out.csv> [env assignments...]
OUT=$1; shift
export MASTER_ADDR=127.0.0.1 MASTER_PORT=29577 WORLD_SIZE=2 NCCL_DEBUG=WARN
for kv in "$@"; do export "$kv"; done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,nvltx__bytes.sum,nvltx__bytes_data_user.sum,nvltx__bytes_data_protocol.sum,nvlrx__bytes.sum,nvlrx__bytes_data_user.sum
RANK=1 LOCAL_RANK=1 timeout 300 python bench.py --gpus 2 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_rank1.log 2>&1 &
P1=$!
RANK=0 LOCAL_RANK=0 timeout 300 ncu --metrics $M --clock-control none -k regex:"fft_kernel|push_kernel" -s 9 -c 6 --csv --log-file $OUT python bench.py --gpus 2 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_rank0.log 2>&1
wait $P1
tail -2 gpurun_out/ncu_rank0.log
