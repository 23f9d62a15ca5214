export NCCL_DEBUG=WARN
N=8
echo "== p2p parity (2x4)"; D2D_P2P=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/mgpu_check.py 2 4 > gpurun_out/mgpu_p2p_$N.txt 2>&1; grep -E "mgpu_check|FAIL|Error" gpurun_out/mgpu_p2p_$N.txt | head
echo "== bench p2p"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N --steps 20 --warmup 5 --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_p2p_$N.json
echo "== bench nccl"; D2D_P2P=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus $N --steps 20 --warmup 5 --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_nccl_$N.json
