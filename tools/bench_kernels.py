"""Per-kernel roofline microbenchmark (device-timed on the library's stream with CUDA events).

usage: python tools/bench_kernels.py [--n 1024] [--prec f64] [--reps 5]
Prints achieved algorithmic GB/s of every 1-D stage on a cube that fills one GPU like the
headline config, and the r2c+c2r pair time.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from __graft_entry__ import package


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1024)
    ap.add_argument("--prec", default="f64")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--shape", default=None, help="nx,ny,nz (default n,n,n)")
    ap.add_argument("--only1d", action="store_true")
    ap.add_argument("--only3d", action="store_true")
    ap.add_argument("--fmt", default="ZX", help="which 3-D pairs: Z, X or ZX")
    ap.add_argument("--axes", default="0,1,2")
    ap.add_argument("--oop", action="store_true", help="1-D stages out of place")
    ap.add_argument("--warm", type=int, default=2)
    a = ap.parse_args()
    p = package()
    shape = tuple(int(x) for x in a.shape.split(",")) if a.shape else (a.n,) * 3
    rdt, cdt = (torch.float64, torch.complex128) if a.prec == "f64" else (torch.float32, torch.complex64)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    d2d = p.decomp_2d_init(*shape, 1, 1)
    d2d.set_blocking(False)
    nx, ny, nz = shape
    cz = None if a.only3d else d2d.alloc_x(cdt)  # (nx,ny,nz) complex
    if cz is not None:
        cz.real.normal_()
    co = d2d.alloc_x(cdt) if a.oop else None
    res = {}
    for axis in ([] if a.only3d else [int(x) for x in a.axes.split(',')]):
        for _ in range(a.warm):
            d2d.c2c_1m(cz, axis, -1, out=co)
        d2d.sync()
        d2d.profile_reset()
        d2d.profile(True)
        for _ in range(a.reps):
            d2d.c2c_1m(cz, axis, -1, out=co)
        d2d.sync()
        d2d.profile(False)
        for k, (ms, calls, by) in d2d.profile_read().items():
            res[k] = (ms / calls, by / calls)
        cz.real.normal_()
        cz.imag.zero_()
    del cz
    torch.cuda.empty_cache()
    # full 3-D pair, PHYSICAL_IN_Z (headline) and PHYSICAL_IN_X
    for fmt, name in (() if a.only1d else [(f, nm) for f, nm in ((p.PHYSICAL_IN_Z, "Z"), (p.PHYSICAL_IN_X, "X")) if nm in a.fmt]):
        eng = p.decomp_2d_fft_init(fmt, dtype=rdt)
        a_in = (d2d.alloc_z if fmt == p.PHYSICAL_IN_Z else d2d.alloc_x)(rdt, eng.ph)
        a_out = (d2d.alloc_x if fmt == p.PHYSICAL_IN_Z else d2d.alloc_z)(cdt, eng.sp)
        a_in.uniform_(-1, 1)
        for _ in range(2):
            eng.fft_3d(a_in, a_out)
            eng.fft_3d(a_out, a_in)
            a_in.mul_(1.0 / (nx * ny * nz))
        d2d.sync()
        d2d.profile_reset()
        d2d.profile(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(a.reps):
            eng.fft_3d(a_in, a_out)
            eng.fft_3d(a_out, a_in)
        d2d.sync()
        e1.record()
        torch.cuda.synchronize()
        d2d.profile(False)
        prof = d2d.profile_read()
        pair_ms = (prof["fft_r2c"][0] + prof["fft_c2r"][0]) / a.reps
        for k, (ms, calls, by) in prof.items():
            if k not in ("fft_r2c", "fft_c2r"):
                res[f"{name}:{k}"] = (ms / calls, by / calls)
        res[f"{name}:pair"] = (pair_ms, 0)
        eng.fin()
        del a_in, a_out
        torch.cuda.empty_cache()
    import math
    N = nx * ny * nz
    flops = 5.0 * N * math.log2(N)
    print(f"shape={shape} prec={a.prec} peak(measured copy)={peak:.0f} GB/s")
    for k, (ms, by) in res.items():
        if by:
            gbs = by / ms / 1e6
            print(f"{k:16s} {ms:9.3f} ms  {gbs:8.1f} GB/s  {gbs / peak * 100:5.1f}% of measured HBM")
        else:
            print(f"{k:16s} {ms:9.3f} ms  pair -> {flops / ms / 1e6:8.1f} GFLOP/s (5 N log2 N)")
    p.decomp_2d_finalize()


if __name__ == "__main__":
    main()
