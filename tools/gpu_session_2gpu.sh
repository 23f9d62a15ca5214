#!/bin/bash
# Two-GPU session of record: multi-process parity tests, the bench line at N = 2, the reference arm under torchrun, NVLink
# counters of the fused producer kernels (rank 0 under ncu) and of the store-pattern microbenchmark.
set -x
mkdir -p gpurun_out
O=gpurun_out
export NCCL_DEBUG=WARN
nvidia-smi --query-gpu=index,name --format=csv
date +%T
timeout 900 python -u -m pytest tests/test_gpu_multiproc.py -m gpu -q -p no:cacheprovider > $O/pytest_gpu_multiproc.log 2>&1; tail -3 $O/pytest_gpu_multiproc.log
date +%T
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29523 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; cat $O/bench_2gpu.json; tail -3 $O/bench_2gpu.err
timeout 300 $TR --master-port 29524 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/bench_ref_2gpu.json 2> $O/bench_ref_2gpu.err; cut -c1-200 $O/bench_ref_2gpu.json
date +%T
bash tools/ncu_rank0_2gpu.sh $O/ncu_nvlink_fused_rank0.csv D2D_EXCHANGE=fused
grep -c fft_kernel $O/ncu_nvlink_fused_rank0.csv
date +%T
M=gpu__time_duration.sum,nvltx__bytes.sum,nvltx__bytes_data_user.sum,nvltx__bytes_data_protocol.sum,dram__bytes_read.sum
for a in "1 128 296 256" "1 512 296 256" "1 64 296 256" "3 32768 32 32"; do timeout 120 ncu --metrics $M --clock-control none -s 2 -c 1 --csv --log-file $O/ncu_nvlbench_$(echo $a | tr ' ' '_').csv tools/micro/nvlbench $a > /dev/null 2>&1; done
date +%T
