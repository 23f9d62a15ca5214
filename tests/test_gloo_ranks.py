"""N > 1 host logic on CPU: two real processes (torch.distributed, gloo, 127.0.0.1) run the private
wire-layout exchange of decomp_2d_fft_3d with the library's own counts / displacements / peer map
(d2d_debug_link_map; host-only, no device) and paired send/recv -- the same structure as the grouped
ncclSend/ncclRecv of csrc/transport_nccl.cpp (reference: src/decomp_2d_nccl.f90:214-473) -- and must
rebuild exactly the destination pencil of the reference transpose (oracle) on every rank."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, shape, grid, link, padq, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import oracle as orc
    from test_wire_layout import AXES, link_map, offsets
    from util import pkg
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p = pkg()
        P, Q = link
        nx, ny, nz = shape
        g = np.arange(nx * ny * nz, dtype=np.float64).reshape(shape, order="F")
        src = orc.scatter(g, grid, P)[rank]
        want = orc.scatter(g, grid, Q)[rank]
        dec = p.DecompInfo.for_rank(nx, ny, nz, grid[0], grid[1], rank)
        prod = link_map(p, dec, P, Q, 0, padq)
        cons = link_map(p, dec, Q, P, 1, padq)
        sz = (dec.xsz, dec.ysz, dec.zsz)[P]
        ax = AXES[P]
        idx = np.indices(sz)
        off, _ = offsets(prod, idx[ax[0]].ravel(), idx[ax[1]].ravel(), idx[ax[2]].ravel())
        sendbuf = torch.full((prod["disp"][-1] + prod["cnt"][-1],), float("nan"), dtype=torch.float64)
        sendbuf[torch.from_numpy(off)] = torch.from_numpy(src[idx[0].ravel(), idx[1].ravel(), idx[2].ravel()])
        recvbuf = torch.full((cons["disp"][-1] + cons["cnt"][-1],), float("nan"), dtype=torch.float64)
        col = (P == 0 or Q == 0)
        c1, c2 = rank // grid[1], rank % grid[1]
        me = c1 if col else c2
        npeers = grid[0] if col else grid[1]
        reqs = []
        for k in range(1, npeers):  # staggered peers, self excluded (the self block never moves)
            m = (me + k) % npeers
            peer = (m * grid[1] + c2) if col else (c1 * grid[1] + m)
            reqs.append(dist.isend(sendbuf[prod["disp"][m]: prod["disp"][m] + prod["cnt"][m]], peer))
            reqs.append(dist.irecv(recvbuf[cons["disp"][m]: cons["disp"][m] + cons["cnt"][m]], peer))
        for r in reqs:
            r.wait()
        sz = (dec.xsz, dec.ysz, dec.zsz)[Q]
        ax = AXES[Q]
        idx = np.indices(sz)
        off, sel = offsets(cons, idx[ax[0]].ravel(), idx[ax[1]].ravel(), idx[ax[2]].ravel())
        sb, rb = sendbuf.numpy(), recvbuf.numpy()
        vals = np.where(sel == 1, sb[np.minimum(off, len(sb) - 1)], rb[np.minimum(off, len(rb) - 1)])
        got = np.zeros(sz)
        got[idx[0].ravel(), idx[1].ravel(), idx[2].ravel()] = vals
        ok = bool(np.array_equal(got, want))
        t = torch.tensor([1 if ok else 0])
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if rank == 0:
            with open(out, "w") as f:
                f.write(str(int(t.item())))
        dec.finalize()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("grid,link", [((1, 2), (2, 1)), ((1, 2), (1, 2)), ((2, 1), (0, 1)), ((2, 1), (1, 0))])
@pytest.mark.parametrize("shape", [(17, 13, 11), (16, 8, 33)])
def test_two_process_exchange_rebuilds_the_reference_pencil(tmp_path, shape, grid, link):
    import torch.multiprocessing as mp
    out = str(tmp_path / "ok.txt")
    mp.spawn(_worker, args=(2, _free_port(), shape, grid, link, 8, out), nprocs=2, join=True)
    assert open(out).read() == "1"
