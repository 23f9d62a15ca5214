"""GPU parity of the arbitrary-length FFT kernel (csrc/fft_any.cuh) against the CPU oracle.

The reference's examples default to grids that are NOT powers of two: 17 x 13 x 11 times small factors
(examples/fft_physical_x/fft_c2c_x.f90:18,39-42; fft_r2c_x.f90, fft_c2c_z.f90, fft_r2c_z.f90 alike) and
examples/fft_multiple_grids uses (nx, ny + 2, nz + 16).  Same tolerances as the power-of-two kernels:
max|delta| / max|ref| <= 1e-12 (fp64), <= 1e-5 (fp32), round trip <= eps * 50 per point.
"""
import numpy as np
import pytest

import oracle as orc
from util import pkg, run_ranks

pytestmark = pytest.mark.gpu

TOL = {"f64": 1e-12, "f32": 1e-5}
LENGTHS = [1, 3, 5, 6, 7, 9, 10, 11, 12, 13, 15, 17, 22, 26, 34, 35, 49, 51, 68, 100, 121, 130, 210, 257, 360, 1000, 1001, 3000]


def _ctx():
    import torch
    p = pkg()
    return p, p.Decomp2d(8, 8, 8, 1, 1, device=torch.cuda.current_device()), torch


def _falloc(torch, shape, dtype):
    n1, n2, n3 = shape
    return torch.zeros((n3, n2, n1), dtype=dtype, device="cuda").permute(2, 1, 0)


def _relerr(a, b):
    return np.max(np.abs(a - b)) / np.max(np.abs(b))


def _max_prime(n):
    m, f = 1, 2
    while f * f <= n:
        while n % f == 0:
            m, n = max(m, f), n // f
        f += 1
    return max(m, n) if n > 1 else m


# The Glassman oracle advances its twiddles by recurrence (src/glassman.f90:103): for a prime factor p its own distance to
# the exact DFT grows with p (measured here: 4e-13 at p = 127, 1.0e-12 at 257, 2e-11 at 1021).  Lengths with a larger prime
# factor are therefore pinned against the exact DFT (numpy's pocketfft = what the reference's FFTW build computes) only.
ORACLE_MAX_PRIME = 127


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("axis", [0, 1, 2])
@pytest.mark.parametrize("n", LENGTHS)
def test_c2c_1m_any_length(n, axis, prec):
    p, d2d, torch = _ctx()
    other = [3, 2] if n >= 1000 else [9, 7]
    shape = other[:]
    shape.insert(axis, n)
    rng = np.random.default_rng(n + axis)
    a = np.asfortranarray(rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape))
    cdt = torch.complex128 if prec == "f64" else torch.complex64
    for isign in (-1, 1):
        t = _falloc(torch, shape, cdt)
        t.copy_(torch.from_numpy(a))
        out = _falloc(torch, shape, cdt)
        d2d.c2c_1m(t, axis, isign, out=out)
        exact = np.fft.fft(a, axis=axis) if isign == -1 else np.fft.ifft(a, axis=axis) * n
        assert _relerr(out.cpu().numpy(), exact) < TOL[prec], (n, axis, isign)
        if _max_prime(n) <= ORACLE_MAX_PRIME:
            ref = orc.c2c_1m(a, axis, isign)
            assert _relerr(out.cpu().numpy(), ref) < TOL[prec], (n, axis, isign)
        d2d.c2c_1m(t, axis, isign)  # in place
        assert _relerr(t.cpu().numpy(), exact) < TOL[prec]
    d2d.finalize()


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("axis", [0, 2])
@pytest.mark.parametrize("batch", [(6, 4), (5, 3), (1, 1)])
@pytest.mark.parametrize("n", [3, 5, 6, 9, 11, 13, 17, 22, 26, 34, 51, 100, 257, 1000, 1001])
def test_r2c_c2r_1m_any_length(n, batch, axis, prec):
    """odd and even non-power-of-two lengths: bins 0..n/2; Im(bin 0) and, for even n, Im(bin n/2) are ignored by c2r
    (src/fft_generic.f90:320-337 takes the real part of a c2c)."""
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
    use_oracle = _max_prime(n) <= ORACLE_MAX_PRIME
    ref = orc.r2c_1m(a, axis) if use_oracle else np.fft.rfft(a, axis=axis)
    assert _relerr(out.cpu().numpy(), ref) < TOL[prec]
    spec = np.asfortranarray(rng.uniform(-1, 1, cshape) + 1j * rng.uniform(-1, 1, cshape))
    tc = _falloc(torch, cshape, cdt)
    tc.copy_(torch.from_numpy(spec))
    back = _falloc(torch, shape, rdt)
    d2d.c2r_1m(tc, back, axis)
    # numpy's irfft drops Im(bin 0) / Im(bin n/2) exactly like the reference's "real part of a c2c"
    ref_r = orc.c2r_1m(spec, n, axis) if use_oracle else np.fft.irfft(spec, n=n, axis=axis) * n
    assert _relerr(back.cpu().numpy(), ref_r) < TOL[prec]
    d2d.finalize()


SHAPES_3D = [(17, 13, 11), (34, 26, 22), (12, 18, 20), (16, 18, 32)]  # the examples' defaults, x2, composite, fft_multiple_grids-like


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("fmt", [orc.PHYSICAL_IN_X, orc.PHYSICAL_IN_Z])
@pytest.mark.parametrize("grid", [(1, 1), (1, 2), (2, 2), (2, 4), (4, 2)])
@pytest.mark.parametrize("shape", SHAPES_3D)
def test_fft_3d_r2c_c2r_any_length(shape, grid, fmt, prec):
    """fft_r2c_x.f90 / fft_r2c_z.f90 at the reference's default (non power-of-two) sizes."""
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
@pytest.mark.parametrize("grid", [(1, 1), (2, 2), (2, 4)])
@pytest.mark.parametrize("shape", SHAPES_3D[:3])
def test_fft_3d_c2c_any_length(shape, grid, fmt, prec):
    """fft_c2c_x.f90 / fft_c2c_z.f90 at the default sizes: forward spectrum vs oracle, round trip."""
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


def test_mixed_lengths_one_power_of_two_axis():
    """one axis on the compiled power-of-two kernels (TMA path at 256), the others on the any-length kernel"""
    import torch
    p = pkg()
    shape, grid = (256, 12, 10), (2, 2)
    rng = np.random.default_rng(2)
    g = np.asfortranarray(rng.uniform(-1, 1, shape))
    ins = orc.scatter(g, grid, 0)
    ref = orc.fft_3d_r2c_world(shape, grid, orc.PHYSICAL_IN_X, ins)

    def body(rank, group):
        d2d = p.Decomp2d(*shape, grid[0], grid[1], rank=rank, nranks=4, group=group, device=0)
        eng = p.Decomp2dFFTEngine(d2d, p.PHYSICAL_IN_X)
        in_r, out_c = d2d.alloc_x(torch.float64, eng.ph), d2d.alloc_z(torch.complex128, eng.sp)
        in_r.copy_(torch.from_numpy(ins[rank]))
        eng.fft_3d(in_r, out_c)
        out = out_c.cpu().numpy()
        eng.fin()
        d2d.finalize()
        return out

    res = run_ranks(4, body)
    smax = max(np.max(np.abs(s)) for s in ref)
    for r in range(4):
        assert np.max(np.abs(res[r] - ref[r])) / smax < 1e-12


@pytest.mark.parametrize("skip", [(True, False, True), (False, True, False), (False, False, True)])
@pytest.mark.parametrize("grid", [(1, 1), (2, 2)])
@pytest.mark.parametrize("shape", [(34, 26, 22), (12, 40, 1021)])
def test_fft_c2c_skip_flags_any_length(shape, grid, skip):
    """fft_c2c_x_skip.f90 (opt_skip_XYZ_c2c) at sizes without a compiled plan, forward and backward: a skipped axis only moves
    its data through the maps (prefetched input, no passes, no index reversal); 1021 takes the chirp-z path."""
    import torch
    p = pkg()
    rng = np.random.default_rng(3)
    g = np.asfortranarray(rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape))
    ins = orc.scatter(g, grid, 0)
    ref = orc.fft_3d_c2c_world(shape, grid, orc.PHYSICAL_IN_X, orc.FORWARD, ins, skip=skip)
    exact = g
    for ax in range(3):
        if not skip[ax]:
            exact = np.fft.fft(exact, axis=ax)
    nranks = grid[0] * grid[1]

    def body(rank, group):
        d2d = p.Decomp2d(*shape, grid[0], grid[1], rank=rank, nranks=nranks, group=group, device=0)
        eng = p.Decomp2dFFTEngine(d2d, p.PHYSICAL_IN_X, opt_skip_XYZ_c2c=skip)
        a_in, a_out = d2d.alloc_x(torch.complex128), d2d.alloc_z(torch.complex128)
        a_in.copy_(torch.from_numpy(ins[rank]))
        eng.fft_3d(a_in, a_out, p.DECOMP_2D_FFT_FORWARD)
        out = a_out.cpu().numpy()
        back = d2d.alloc_x(torch.complex128)
        eng.fft_3d(a_out, back, p.DECOMP_2D_FFT_BACKWARD)
        res = (out, back.cpu().numpy())
        eng.fin()
        d2d.finalize()
        return res

    res = run_ranks(nranks, body)
    spec = orc.gather([x[0] for x in res], shape, grid, 2)
    assert _relerr(spec, exact) < 1e-12
    if _max_prime(max(shape)) <= ORACLE_MAX_PRIME:
        for r in range(nranks):
            assert _relerr(res[r][0], ref[r]) < 1e-12
    scale = np.prod([shape[ax] for ax in range(3) if not skip[ax]])
    rt = orc.gather([x[1] for x in res], shape, grid, 0) / scale
    assert np.max(np.abs(rt - g)) < 1e-12
