"""Run one r2c + c2r pair (for ncu captures).  usage: python tools/run_pair.py nx ny nz [fmt z|x] [prec] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from __graft_entry__ import package
p = package()
nx, ny, nz = [int(v) for v in sys.argv[1:4]]
fmt = p.PHYSICAL_IN_Z if (len(sys.argv) < 5 or sys.argv[4] == "z") else p.PHYSICAL_IN_X
prec = sys.argv[5] if len(sys.argv) > 5 else "f64"
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 1
rdt, cdt = (torch.float64, torch.complex128) if prec == "f64" else (torch.float32, torch.complex64)
d2d = p.decomp_2d_init(nx, ny, nz, 1, 1)
eng = p.decomp_2d_fft_init(fmt, dtype=rdt)
a_in = (d2d.alloc_z if fmt == p.PHYSICAL_IN_Z else d2d.alloc_x)(rdt, eng.ph)
a_out = (d2d.alloc_x if fmt == p.PHYSICAL_IN_Z else d2d.alloc_z)(cdt, eng.sp)
a_in.uniform_(-1, 1)
for _ in range(reps):
    eng.fft_3d(a_in, a_out)
    eng.fft_3d(a_out, a_in)
torch.cuda.synchronize()
print("done")
p.decomp_2d_finalize()
