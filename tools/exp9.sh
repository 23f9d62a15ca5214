export NCCL_DEBUG=WARN
N=${1:-4}
echo "== p2p parity"; D2D_P2P=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/mgpu_check.py > gpurun_out/mgpu_p2p_$N.txt 2>&1; grep -E "mgpu_check|FAIL|Error" gpurun_out/mgpu_p2p_$N.txt | head
echo "== nccl parity"; D2D_P2P=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 tools/mgpu_check.py > gpurun_out/mgpu_nccl_$N.txt 2>&1; grep -E "mgpu_check|FAIL|Error" gpurun_out/mgpu_nccl_$N.txt | head
echo "== bench p2p"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_p2p_$N.json
echo "== bench nccl"; D2D_P2P=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_nccl_$N.json
