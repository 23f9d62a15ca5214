"""GPU parity of the compiled 3 * 2^k and 5 * 2^k kernels (radix-24 / radix-20 plans of csrc/fft_kernel.cuh on the
cp.async and TMA-staged kernels) against the CPU oracle, through the C ABI.

The reference takes these lengths through cuFFT's mixed-radix plans (src/fft_cufft.f90:73-258) and, in its generic
backend, through the factor loop of SPCFFT (src/glassman.f90:29-67); typical grids of its users (192, 384, 768, 640, 1280
points per direction) are of this kind.  Same tolerances as everywhere: max|delta| / max|ref| <= 1e-12 (fp64),
<= 1e-5 (fp32, against the fp64 oracle), round trip <= eps * 50 per point.
"""
import numpy as np
import pytest

import oracle as orc
from util import pkg, run_ranks

pytestmark = pytest.mark.gpu

TOL = {"f64": 1e-12, "f32": 1e-5}
MIXED = [6, 12, 24, 48, 96, 192, 384, 768, 1536, 3072, 10, 20, 40, 80, 160, 320, 640, 1280]


def _ctx():
    import torch
    p = pkg()
    return p, p.Decomp2d(8, 8, 8, 1, 1, device=torch.cuda.current_device()), torch


def _falloc(torch, shape, dtype):
    n1, n2, n3 = shape
    return torch.zeros((n3, n2, n1), dtype=dtype, device="cuda").permute(2, 1, 0)


def _relerr(a, b):
    return np.max(np.abs(a - b)) / np.max(np.abs(b))


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("axis", [0, 1, 2])
@pytest.mark.parametrize("n", MIXED)
def test_c2c_1m_mixed_radix(n, axis, prec):
    """odd batch extents (partial tiles, cp.async kernels for fp32) and even ones (16-byte strides: the TMA kernels)"""
    p, d2d, torch = _ctx()
    cdt = torch.complex128 if prec == "f64" else torch.complex64
    for other in ([5, 3], [8, 6]) if n >= 768 else ([9, 7], [16, 6]):
        shape = other[:]
        shape.insert(axis, n)
        rng = np.random.default_rng(n + axis)
        a = np.asfortranarray(rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape))
        for isign in (-1, 1):
            t = _falloc(torch, shape, cdt)
            t.copy_(torch.from_numpy(a))
            out = _falloc(torch, shape, cdt)
            d2d.c2c_1m(t, axis, isign, out=out)
            ref = orc.c2c_1m(a, axis, isign)
            exact = np.fft.fft(a, axis=axis) if isign == -1 else np.fft.ifft(a, axis=axis) * n
            assert _relerr(out.cpu().numpy(), ref) < TOL[prec], (n, axis, isign, other)
            assert _relerr(out.cpu().numpy(), exact) < TOL[prec], (n, axis, isign, other)
            assert np.array_equal(t.cpu().numpy(), a.astype(t.cpu().numpy().dtype)), "input must be preserved when out != in"
            d2d.c2c_1m(t, axis, isign)  # in place
            assert _relerr(t.cpu().numpy(), ref) < TOL[prec]
    d2d.finalize()


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("axis", [0, 2])
@pytest.mark.parametrize("batch", [(6, 4), (5, 3), (1, 1), (16, 2)])
@pytest.mark.parametrize("n", MIXED)
def test_r2c_c2r_1m_mixed_radix(n, batch, axis, prec):
    p, d2d, torch = _ctx()
    shape = list(batch)
    shape.insert(axis, n)
    rng = np.random.default_rng(n)
    a = np.asfortranarray(rng.uniform(-1, 1, shape))
    rdt, cdt = (torch.float64, torch.complex128) if prec == "f64" else (torch.float32, torch.complex64)
    cshape = list(shape)
    cshape[axis] = n // 2 + 1
    t = _falloc(torch, shape, rdt)
    t.copy_(torch.from_numpy(a))
    out = _falloc(torch, cshape, cdt)
    d2d.r2c_1m(t, out, axis)
    ref = orc.r2c_1m(a, axis)
    assert _relerr(out.cpu().numpy(), ref) < TOL[prec]
    assert _relerr(out.cpu().numpy(), np.fft.rfft(a, axis=axis)) < TOL[prec]
    # c2r on a spectrum whose DC / Nyquist bins carry imaginary parts (reference semantics: ignored)
    spec = np.asfortranarray(rng.uniform(-1, 1, cshape) + 1j * rng.uniform(-1, 1, cshape))
    tc = _falloc(torch, cshape, cdt)
    tc.copy_(torch.from_numpy(spec))
    back = _falloc(torch, shape, rdt)
    d2d.c2r_1m(tc, back, axis)
    assert _relerr(back.cpu().numpy(), orc.c2r_1m(spec, n, axis)) < TOL[prec]
    d2d.finalize()


# (shape, grids): small grids on every process grid, larger ones (TMA kernels inside the chains, ragged spectral pencils:
# 320/2+1 = 161, 40/2+1 = 21) on fewer
SHAPES_3D = [((96, 80, 48), [(1, 1), (1, 2), (2, 2), (2, 4), (4, 2)]),
             ((320, 384, 256), [(1, 1), (2, 2)]),
             ((768, 24, 640), [(1, 1), (2, 4)]),
             ((40, 24, 768), [(1, 1), (2, 2)])]
CASES_3D = [(s, g) for s, gs in SHAPES_3D for g in gs]


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("fmt", [orc.PHYSICAL_IN_X, orc.PHYSICAL_IN_Z])
@pytest.mark.parametrize("shape,grid", CASES_3D)
def test_fft_3d_r2c_c2r_mixed_radix(shape, grid, fmt, prec):
    import torch
    p = pkg()
    rdt = np.float64 if prec == "f64" else np.float32
    trd, tcd = (torch.float64, torch.complex128) if prec == "f64" else (torch.float32, torch.complex64)
    rng = np.random.default_rng(7)
    g = np.asfortranarray(rng.uniform(-1, 1, shape)).astype(rdt)
    pin = 0 if fmt == orc.PHYSICAL_IN_X else 2
    ins = orc.scatter(g, grid, pin)
    ref_spec = orc.fft_3d_r2c_world(shape, grid, fmt, [a.astype(np.float64) for a in ins])
    nranks = grid[0] * grid[1]

    def body(rank, group):
        d2d = p.Decomp2d(*shape, grid[0], grid[1], rank=rank, nranks=nranks, group=group, device=0)
        eng = p.Decomp2dFFTEngine(d2d, fmt, dtype=trd)
        alloc_in = d2d.alloc_x if fmt == orc.PHYSICAL_IN_X else d2d.alloc_z
        alloc_out = d2d.alloc_z if fmt == orc.PHYSICAL_IN_X else d2d.alloc_x
        in_r, out_c = alloc_in(trd, eng.ph), alloc_out(tcd, eng.sp)
        in_r.copy_(torch.from_numpy(ins[rank]))
        eng.fft_3d(in_r, out_c)
        assert np.array_equal(in_r.cpu().numpy(), ins[rank]), "r2c must not modify its input"
        spec = out_c.cpu().numpy()
        rt = alloc_in(trd, eng.ph)
        eng.fft_3d(out_c, rt)
        res = (spec, rt.cpu().numpy())
        eng.fin()
        d2d.finalize()
        return res

    res = run_ranks(nranks, body)
    smax = max(np.max(np.abs(s)) for s in ref_spec if s.size)
    for r in range(nranks):
        if ref_spec[r].size:
            assert np.max(np.abs(res[r][0] - ref_spec[r])) / smax < TOL[prec], ("spectrum", r)
    rt = orc.gather([x[1] for x in res], shape, grid, pin).astype(np.float64) / np.prod(shape)
    assert np.sum(np.abs(rt - g)) / np.prod(shape) < np.finfo(rdt).eps * 50


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("fmt", [orc.PHYSICAL_IN_X, orc.PHYSICAL_IN_Z])
@pytest.mark.parametrize("shape,grid", [c for c in CASES_3D if c[1] in ((1, 1), (2, 2), (2, 4))])
def test_fft_3d_c2c_mixed_radix(shape, grid, fmt, prec):
    import torch
    p = pkg()
    rdt, cdt = (np.float64, np.complex128) if prec == "f64" else (np.float32, np.complex64)
    tcd = torch.complex128 if prec == "f64" else torch.complex64
    rng = np.random.default_rng(11)
    g = np.asfortranarray(rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)).astype(cdt)
    pin, pout = (0, 2) if fmt == orc.PHYSICAL_IN_X else (2, 0)
    ins = orc.scatter(g, grid, pin)
    ref_spec = orc.fft_3d_c2c_world(shape, grid, fmt, orc.FORWARD, [a.astype(np.complex128) for a in ins])
    nranks = grid[0] * grid[1]

    def body(rank, group):
        d2d = p.Decomp2d(*shape, grid[0], grid[1], rank=rank, nranks=nranks, group=group, device=0)
        eng = p.Decomp2dFFTEngine(d2d, fmt, dtype=torch.float64 if prec == "f64" else torch.float32)
        a_in = (d2d.alloc_x if pin == 0 else d2d.alloc_z)(tcd)
        a_out = (d2d.alloc_x if pout == 0 else d2d.alloc_z)(tcd)
        a_in.copy_(torch.from_numpy(ins[rank]))
        eng.fft_3d(a_in, a_out, p.DECOMP_2D_FFT_FORWARD)
        assert np.array_equal(a_in.cpu().numpy(), ins[rank]), "c2c (not inplace) must preserve its input"
        spec = a_out.cpu().numpy()
        a_back = (d2d.alloc_x if pin == 0 else d2d.alloc_z)(tcd)
        eng.fft_3d(a_out, a_back, p.DECOMP_2D_FFT_BACKWARD)
        res = (spec, a_back.cpu().numpy())
        eng.fin()
        d2d.finalize()
        return res

    res = run_ranks(nranks, body)
    smax = max(np.max(np.abs(s)) for s in ref_spec if s.size)
    for r in range(nranks):
        if ref_spec[r].size:
            assert np.max(np.abs(res[r][0] - ref_spec[r])) / smax < TOL[prec]
    rt = orc.gather([x[1] for x in res], shape, grid, pin) / np.prod(shape)
    assert np.sum(np.abs(rt - g)) / np.prod(shape) < np.finfo(rdt).eps * 50
