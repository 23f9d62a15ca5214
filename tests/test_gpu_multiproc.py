"""One rank per GPU over NCCL (torchrun, 127.0.0.1): parity of the fused peer-store path and of the NCCL exchange path
against the oracle.  Needs >= 2 GPUs (skipped on a single-GPU box; the thread-per-rank tests of test_gpu_fft3d.py
cover the same maps on one device)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("p2p", ["1", "0"])
def test_torchrun_parity(p2p):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 4 if n >= 4 else 2
    env = dict(os.environ, D2D_P2P=p2p)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533" if p2p == "1" else "29534", os.path.join(ROOT, "tools", "mgpu_check.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "failures 0" in r.stdout
