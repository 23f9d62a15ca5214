import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np, torch
import oracle as orc
from util import pkg, run_ranks
p = pkg()
for shape, grid in (((32, 16, 64), (1, 2)), ((32, 16, 64), (2, 1)), ((32,16,64),(2,2))):
  for fmt in (3, 1):
    rng = np.random.default_rng(1)
    g = np.asfortranarray(rng.uniform(-1, 1, shape))
    pin, pout = (0, 2) if fmt == 1 else (2, 0)
    ins = orc.scatter(g, grid, pin)
    ref = orc.fft_3d_r2c_world(shape, grid, fmt, ins)
    refb = orc.fft_3d_c2r_world(shape, grid, fmt, ref)
    n = grid[0] * grid[1]
    def body(rank, group):
        d2d = p.Decomp2d(*shape, grid[0], grid[1], rank=rank, nranks=n, group=group, device=0)
        eng = p.Decomp2dFFTEngine(d2d, fmt)
        a_in = (d2d.alloc_x if pin == 0 else d2d.alloc_z)(torch.float64, eng.ph)
        a_out = (d2d.alloc_x if pout == 0 else d2d.alloc_z)(torch.complex128, eng.sp)
        a_in.copy_(torch.from_numpy(ins[rank]))
        eng.fft_3d(a_in, a_out)
        o = a_out.cpu().numpy()
        a_out.copy_(torch.from_numpy(ref[rank]))
        a_in.zero_()
        eng.fft_3d(a_out, a_in)
        ob = a_in.cpu().numpy()
        eng.fin(); d2d.finalize(); return o, ob
    res = run_ranks(n, body)
    for r in range(n):
        for which, (got, want) in enumerate(((res[r][0], ref[r]), (res[r][1], refb[r]))):
            err = np.abs(got - want)
            bad = np.argwhere(err > 1e-9 * np.max(np.abs(want)))
            print(f"grid={grid} fmt={fmt} {'r2c' if which == 0 else 'c2r'} rank={r} shape={got.shape} maxerr={err.max():.3e} nbad={len(bad)}/{got.size}",
                  "i:", (bad[:, 0].min(), bad[:, 0].max()) if len(bad) else "", "j:", (bad[:, 1].min(), bad[:, 1].max()) if len(bad) else "",
                  "k:", (bad[:, 2].min(), bad[:, 2].max()) if len(bad) else "", "first:", bad[:4].tolist() if len(bad) else "")
