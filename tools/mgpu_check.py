"""Multi-process (one rank per GPU, NCCL bootstrap) parity check of the 3-D transforms and the bare transposes
against the CPU oracle.  Run under torchrun:
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/mgpu_check.py [p_row p_col]
Every rank evaluates the oracle for the whole (small) world and compares its own pencils.  Exit code 0 = parity."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
import torch.distributed as dist

import oracle as orc
from __graft_entry__ import package


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    # MGPU_BACKEND=gloo: torch.distributed only carries the bootstrap all-gather (transport "boot": no NCCL anywhere), which
    # also lets several ranks share one GPU -- the peer-memory exchange (CUDA IPC + copy engines + stream-ordered flags) is the
    # same code whether the peer's buffer lives on another GPU or on this one
    backend = os.environ.get("MGPU_BACKEND", "nccl")
    dev = local % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    if backend == "nccl":
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    else:
        dist.init_process_group("gloo")
    p = package()
    grids = {1: [(1, 1)], 2: [(1, 2), (2, 1)], 4: [(2, 2), (1, 4), (4, 1)], 8: [(2, 4), (4, 2)]}[world]
    if len(sys.argv) > 2:
        grids = [(int(sys.argv[1]), int(sys.argv[2]))]
    # powers of two (cp.async and TMA kernels), 3 * 2^k / 5 * 2^k (compiled mixed-radix plans), the reference's default
    # 17 x 13 x 11 type (any-length kernel with prefetch)
    shapes = ((32, 16, 64), (256, 64, 512), (64, 256, 32), (384, 40, 320), (34, 26, 22))
    if os.environ.get("MGPU_SHAPES") == "small":
        shapes = ((32, 16, 64), (64, 256, 32), (34, 26, 22))
    worst = 0.0
    nfail = 0
    for grid in grids:
        for shape in shapes:
            for fmt in (p.PHYSICAL_IN_Z, p.PHYSICAL_IN_X):
                if fmt == p.PHYSICAL_IN_X and shape[0] % 2:
                    continue
                d2d = p.decomp_2d_init_from_torch_distributed(*shape, *grid)
                eng = p.decomp_2d_fft_init(fmt)
                rng = np.random.default_rng(5)
                g = np.asfortranarray(rng.uniform(-1, 1, shape))
                pin, pout = (0, 2) if fmt == p.PHYSICAL_IN_X else (2, 0)
                ins = orc.scatter(g, grid, pin)
                ref = orc.fft_3d_r2c_world(shape, grid, fmt, ins)
                refb = orc.fft_3d_c2r_world(shape, grid, fmt, ref)
                a_in = (d2d.alloc_x if pin == 0 else d2d.alloc_z)(torch.float64, eng.ph)
                a_out = (d2d.alloc_x if pout == 0 else d2d.alloc_z)(torch.complex128, eng.sp)
                a_in.copy_(torch.from_numpy(ins[rank]))
                for rep in range(3):  # repeated calls exercise the flag epochs / buffer reuse
                    eng.fft_3d(a_in, a_out)
                smax = max(np.max(np.abs(s)) for s in ref)
                e1 = float(np.max(np.abs(a_out.cpu().numpy() - ref[rank])) / smax) if ref[rank].size else 0.0
                back = (d2d.alloc_x if pin == 0 else d2d.alloc_z)(torch.float64, eng.ph)
                for rep in range(2):
                    eng.fft_3d(a_out, back)
                bmax = max(np.max(np.abs(s)) for s in refb)
                e2 = float(np.max(np.abs(back.cpu().numpy() - refb[rank])) / bmax) if refb[rank].size else 0.0
                # c2c forward on the same grid
                gc = np.asfortranarray(rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape))
                cin = orc.scatter(gc, grid, pin)
                refc = orc.fft_3d_c2c_world(shape, grid, fmt, orc.FORWARD, cin)
                c_in = (d2d.alloc_x if pin == 0 else d2d.alloc_z)(torch.complex128, eng.ph)
                c_out = (d2d.alloc_x if pout == 0 else d2d.alloc_z)(torch.complex128, eng.ph)
                c_in.copy_(torch.from_numpy(cin[rank]))
                eng.fft_3d(c_in, c_out, p.DECOMP_2D_FFT_FORWARD)
                cmax = max(np.max(np.abs(s)) for s in refc)
                e3 = float(np.max(np.abs(c_out.cpu().numpy() - refc[rank])) / cmax) if refc[rank].size else 0.0
                # bare transposes, bit-exact (test2d)
                idx = np.arange(np.prod(shape), dtype=np.float64).reshape(shape, order="F")
                want = [orc.scatter(idx, grid, k) for k in range(3)]
                u1, u2, u3 = d2d.alloc_x(torch.float64), d2d.alloc_y(torch.float64), d2d.alloc_z(torch.float64)
                u1.copy_(torch.from_numpy(want[0][rank]))
                d2d.transpose_x_to_y(u1, u2)
                d2d.transpose_y_to_z(u2, u3)
                t_ok = np.array_equal(u2.cpu().numpy(), want[1][rank]) and np.array_equal(u3.cpu().numpy(), want[2][rank])
                u2.zero_(); u1.zero_()
                d2d.transpose_z_to_y(u3, u2)
                d2d.transpose_y_to_x(u2, u1)
                t_ok = t_ok and np.array_equal(u2.cpu().numpy(), want[1][rank]) and np.array_equal(u1.cpu().numpy(), want[0][rank])
                # halo exchange over the same transport (periodic in x and z), bit-exact
                for pen in range(3):
                    hw = orc.update_halo_world(idx, grid, pen, 2, (True, False, True))
                    hin = (d2d.alloc_x, d2d.alloc_y, d2d.alloc_z)[pen](torch.float64)
                    hin.copy_(torch.from_numpy(want[pen][rank]))
                    d2d.periodic_bc = (True, False, True)
                    hout = d2d.update_halo(hin, 2, opt_pencil=pen + 1)
                    t_ok = t_ok and np.array_equal(hout.cpu().numpy(), hw[rank])
                err = max(e1, e2, e3)
                worst = max(worst, err)
                bad = (err > 1e-12) or not t_ok
                nfail += int(bad)
                if rank == 0 or bad:
                    print(f"rank {rank} grid {grid} shape {shape} fmt {fmt}: r2c {e1:.2e} c2r {e2:.2e} c2c {e3:.2e} transposes "
                          f"{'exact' if t_ok else 'WRONG'} {'FAIL' if bad else 'ok'}", flush=True)
                p.decomp_2d_finalize()
    t = torch.tensor([nfail], device="cuda" if backend == "nccl" else "cpu")
    dist.all_reduce(t)
    if rank == 0:
        print(f"mgpu_check: world {world}, failures {int(t.item())}, worst rel err {worst:.2e}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if int(t.item()) else 0)


if __name__ == "__main__":
    main()
