export NCCL_DEBUG=WARN
D2D_P2P=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/mgpu_check.py > gpurun_out/mgpu_p2p.txt 2>&1
grep -v "^\*\|OMP_NUM" gpurun_out/mgpu_p2p.txt | head -60
