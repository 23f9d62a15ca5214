"""Bare transpose_* timing on ONE GPU with p_row x p_col rank-threads (d2d_group transport: the exchange is a device-to-device
copy, so this measures pack + local exchange + unpack, not NVLink).  usage: python tools/bench_transpose.py [n] [p_row] [p_col]
Prints per-direction ms (rank 0's CUDA-event timers) and the aggregate GB/s over all ranks: each transpose reads and writes
every pencil once in pack, once in the exchange and once in unpack where those steps exist (algorithmic: 2 x pencil bytes)."""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from __graft_entry__ import package


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    p_row = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    p_col = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    reps = 5
    p = package()
    nranks = p_row * p_col
    group = p.Group(nranks)
    out = [None] * nranks

    def body(rank):
        d2d = p.Decomp2d(n, n, n, p_row, p_col, rank=rank, nranks=nranks, group=group, device=0)
        u1, u2, u3 = d2d.alloc_x(torch.complex128), d2d.alloc_y(torch.complex128), d2d.alloc_z(torch.complex128)
        u1.real.normal_()
        for it in range(reps + 1):
            if it == 1:
                d2d.profile_reset()
                d2d.profile(True)
                t0 = time.perf_counter()
            d2d.transpose_x_to_y(u1, u2)
            d2d.transpose_y_to_z(u2, u3)
            d2d.transpose_z_to_y(u3, u2)
            d2d.transpose_y_to_x(u2, u1)
        d2d.sync()
        wall = (time.perf_counter() - t0) / reps
        d2d.profile(False)
        out[rank] = (d2d.profile_read(), wall, u1.numel() * 16)
        d2d.finalize()

    ts = [threading.Thread(target=body, args=(r,)) for r in range(nranks)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    prof, wall, pencil_bytes = out[0]
    total_bytes = 2.0 * pencil_bytes * nranks  # read + write of the whole field per transpose
    print(f"n={n} grid {p_row}x{p_col} complex128, {nranks} rank-threads on one GPU; wall per 4 transposes {wall * 1e3:.2f} ms")
    for k in ("transp_x_y", "transp_y_z", "transp_z_y", "transp_y_x"):
        if k in prof:
            ms = prof[k][0] / prof[k][1]
            print(f"{k:12s} {ms:8.3f} ms on rank 0 (ranks run concurrently)  -> {total_bytes / ms / 1e6:8.1f} GB/s aggregate (2 x field bytes / time)")


if __name__ == "__main__":
    main()
