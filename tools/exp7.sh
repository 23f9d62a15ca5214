export NCCL_DEBUG=WARN
echo "== p2p parity"; D2D_P2P=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/mgpu_check.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -30
echo "== nccl parity"; D2D_P2P=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 tools/mgpu_check.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -4
echo "== bench p2p"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e 2>&1 | tail -1
echo "== bench nccl"; D2D_P2P=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e 2>&1 | tail -1
