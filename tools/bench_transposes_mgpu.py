"""The reference's examples/test2d/timing2d_complex.f90:89-233 restated for one rank per GPU (torchrun): the index-encoded complex
field (m, m-1) goes x -> y -> z -> y -> x `reps` times, exact equality is checked after the timed loop, and the per-direction
device times (CUDA events on the library's stream, max over ranks) are printed with what they mean for the hardware:
   pack / unpack  GB/s of pencil bytes read + written by the copy kernels, against the measured HBM copy rate
   exchange       GB/s of bytes this rank SENDS (self block excluded), against NVLink 900 GB/s
usage: python -m torch.distributed.run --nproc-per-node N tools/bench_transposes_mgpu.py [n] [p_row p_col] [reps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

from __graft_entry__ import package


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    p = package()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    grid = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else {2: (1, 2), 4: (2, 2), 8: (2, 4)}[world]
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
    d2d = p.decomp_2d_init_from_torch_distributed(n, n, n, *grid)
    d2d.set_blocking(False)
    dm = d2d.decomp_main
    u1, u2, u3 = d2d.alloc_x(torch.complex128), d2d.alloc_y(torch.complex128), d2d.alloc_z(torch.complex128)
    # timing2d_complex.f90:100-112: m = i + (j-1) nx + (k-1) nx ny (global, 1-based), data = (m, m-1)
    ax = [torch.arange(dm.xst[d], dm.xst[d] + dm.xsz[d], device=u1.device, dtype=torch.float64) for d in range(3)]
    m = ax[0][:, None, None] + (ax[1][None, :, None] - 1) * n + (ax[2][None, None, :] - 1) * float(n) * n
    u1.copy_(torch.complex(m, m - 1))
    ref = u1.clone()

    def cycle():
        d2d.transpose_x_to_y(u1, u2)
        d2d.transpose_y_to_z(u2, u3)
        d2d.transpose_z_to_y(u3, u2)
        d2d.transpose_y_to_x(u2, u1)

    for _ in range(2):
        cycle()
    d2d.sync()
    dist.barrier()
    d2d.profile_reset()
    d2d.profile(True)
    for _ in range(reps):
        cycle()
    d2d.sync()
    d2d.profile(False)
    ok = bool(torch.equal(u1, ref))
    prof = d2d.profile_read()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    keys = sorted(prof)
    vals = torch.tensor([[prof[k][0] / max(prof[k][1], 1), prof[k][2] / max(prof[k][1], 1)] for k in keys], dtype=torch.float64, device=u1.device)
    mx = vals.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    okt = torch.tensor([int(ok)], device=u1.device)
    dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"timing2d_complex restated: {n}^3 complex128, grid {grid[0]}x{grid[1]}, {reps} cycles, exact round trip: {bool(okt.item())}")
        for i, k in enumerate(keys):
            ms, by = float(mx[i, 0]), float(vals[i, 1])
            if k in ("pack", "unpack"):
                print(f"  {k:11s} {ms:8.3f} ms per launch  {by / ms / 1e6:8.1f} GB/s  = {by / ms / 1e6 / hbm * 100:5.1f} % of the measured HBM copy rate")
            elif k.startswith("a2a_"):
                print(f"  {k:11s} {ms:8.3f} ms  sends {by / 1e6:8.1f} MB -> {by / ms / 1e6:7.1f} GB/s = {by / ms / 1e6 / 900 * 100:5.1f} % of NVLink 900 GB/s")
            elif k.startswith("transp_"):
                print(f"  {k:11s} {ms:8.3f} ms per transpose (pack + exchange + unpack)")
    p.decomp_2d_finalize()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if okt.item() else 1)


if __name__ == "__main__":
    main()
