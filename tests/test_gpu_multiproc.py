"""One rank per GPU over NCCL (torchrun, 127.0.0.1): parity of the fused peer-store path and of the NCCL exchange path
against the oracle.  Needs >= 2 GPUs (skipped on a single-GPU box; the thread-per-rank tests of test_gpu_fft3d.py
cover the same maps on one device)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("p2p", ["1", "0", "fused"])
def test_torchrun_parity(p2p):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 4 if n >= 4 else 2
    env = dict(os.environ, D2D_P2P="0" if p2p == "0" else "1", D2D_FUSED="1" if p2p == "fused" else "0")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", {"1": "29533", "0": "29534", "fused": "29535"}[p2p], os.path.join(ROOT, "tools", "mgpu_check.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "failures 0" in r.stdout


@pytest.mark.parametrize("world,chunks,exchange", [(2, "4", "auto"), (4, "3", "auto"), (2, "1", "pipe"), (4, "4", "fused"), (8, "4", "auto")])
def test_torchrun_parity_shared_gpu(world, chunks, exchange):
    """The default multi-rank data plane -- chunk-pipelined chains and transposes whose blocks are pushed by the copy engines
    into the peers' CUDA-IPC mapped buffers, ordered by stream memory operations -- with `world` processes SHARING GPU 0
    (bootstrap over gloo, no NCCL).  Runs on a single-GPU box: forward spectra, round trips and bit-exact transposes of
    every rank are compared with the oracle by tools/mgpu_check.py."""
    # exchange: auto = pipelined chains (copy-engine pushes) up to 2 ranks per communicator, fused peer stores beyond (the 2 x 4
    # grid of 8 ranks: both links of a chain are real); "fused" on 4 ranks forces the fused chain onto the 2 x 2 grid
    env = dict(os.environ, MGPU_BACKEND="gloo", D2D_TRANSPORT="boot", D2D_CHUNKS=chunks, D2D_EXCHANGE=exchange, MGPU_SHAPES="small",
               CUDA_VISIBLE_DEVICES="0")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29540 + world + int(chunks) + len(exchange)), os.path.join(ROOT, "tools", "mgpu_check.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "failures 0" in r.stdout
