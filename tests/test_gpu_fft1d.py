"""GPU parity of the batched 1-D FFT kernels (through the C ABI) against the CPU oracle.

Tolerances (BASELINE.json north_star): max|delta| / max|ref| <= 1e-12 in fp64 and <= 1e-5 in fp32.
fp32 is compared with the fp64 oracle (the fp32 Glassman oracle itself is ~5e-6 away from the true
DFT per pass); the distance to the fp32 oracle is checked with a looser bound.
"""
import numpy as np
import pytest

import oracle as orc
from util import pkg

pytestmark = pytest.mark.gpu

TOL = {"f64": 1e-12, "f32": 1e-5}


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
@pytest.mark.parametrize("n", [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192])
def test_c2c_1m_vs_oracle(n, axis, prec):
    p, d2d, torch = _ctx()
    other = [5, 3] if n >= 1024 else [9, 7]  # odd batch extents: partial tiles
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
        ref = orc.c2c_1m(a, axis, isign)
        assert _relerr(out.cpu().numpy(), ref) < TOL[prec], (n, axis, isign)
        assert np.array_equal(t.cpu().numpy(), a.astype(t.cpu().numpy().dtype)), "input must be preserved when out != in"
        d2d.c2c_1m(t, axis, isign)  # in place
        assert _relerr(t.cpu().numpy(), ref) < TOL[prec]
    d2d.finalize()


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("axis", [0, 2])
@pytest.mark.parametrize("batch", [(6, 4), (5, 3), (1, 1), (7, 2)])
@pytest.mark.parametrize("n", [2, 4, 8, 16, 64, 256, 512, 1024, 2048])
def test_r2c_c2r_1m_vs_oracle(n, batch, axis, prec):
    """r2c_1m_x/z and c2r_1m_x/z; odd batch extents exercise the unpaired last line and the
    non-vector (unaligned) pair path of the two-for-one real transform."""
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
    # c2r on a spectrum whose DC / Nyquist bins carry imaginary parts (reference semantics: ignored)
    spec = np.asfortranarray(rng.uniform(-1, 1, cshape) + 1j * rng.uniform(-1, 1, cshape))
    tc = _falloc(torch, cshape, cdt)
    tc.copy_(torch.from_numpy(spec))
    back = _falloc(torch, shape, rdt)
    d2d.c2r_1m(tc, back, axis)
    ref_r = orc.c2r_1m(spec, n, axis)
    assert _relerr(back.cpu().numpy(), ref_r) < TOL[prec]
    d2d.finalize()


def test_fp32_distance_to_fp32_oracle_is_reported():
    p, d2d, torch = _ctx()
    n, shape = 1024, (1024, 4, 4)
    rng = np.random.default_rng(0)
    a = np.asfortranarray(rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)).astype(np.complex64)
    t = _falloc(torch, shape, torch.complex64)
    t.copy_(torch.from_numpy(a))
    d2d.c2c_1m(t, 0, -1)
    ref32 = orc.c2c_1m(a, 0, -1)
    ref64 = orc.c2c_1m(a.astype(np.complex128), 0, -1)
    ours, theirs = _relerr(t.cpu().numpy(), ref64), _relerr(ref32, ref64)
    print(f"fp32 n={n}: ours vs fp64 oracle {ours:.2e}; fp32 oracle vs fp64 oracle {theirs:.2e}")
    assert ours < 1e-5 and ours <= theirs * 1.5
    d2d.finalize()


def test_unsupported_length_is_an_error_not_a_fallback():
    """every length up to 2^24 has a kernel (compiled plan, shared-memory mixed radix, two-kernel split, chirp-z); beyond that
    a length with a large prime factor is rejected with an error, never handed to a CPU path"""
    p, d2d, torch = _ctx()
    n = (1 << 24) + 1  # 97 * 257 * 673
    t = _falloc(torch, (n, 1, 1), torch.complex128)
    with pytest.raises(p.Decomp2dError, match="not supported"):
        d2d.c2c_1m(t, 0, -1)
    d2d.finalize()


@pytest.mark.parametrize("n,prec", [(4800, "f64"), (15015, "f64"), (16384, "f64"), (6000, "f64"), (10000, "f32")])
@pytest.mark.parametrize("axis", [0, 1, 2])
def test_long_lines_two_kernel_split(n, prec, axis):
    """lengths beyond the shared-memory kernel (fp64: 4265, fp32: 8532) run as two launches through a scratch array
    (n = n1 n2, four-step split): c2c both directions, r2c and c2r.  Pinned against the exact DFT (numpy's pocketfft): the
    oracle -- like the reference's generic backend, src/glassman.f90:29-67 -- takes any length, but its recurrence twiddles
    drift to ~5e-12 at n = 15015, so it is only required to agree to 1e-10."""
    p, d2d, torch = _ctx()
    cdt, rdt = (np.complex128, np.float64) if prec == "f64" else (np.complex64, np.float32)
    tcd, trd = (torch.complex128, torch.float64) if prec == "f64" else (torch.complex64, torch.float32)
    tol = 1e-12 if prec == "f64" else 1e-5
    shape = [3, 4, 2]
    shape[axis] = n
    shape = tuple(shape)
    rng = np.random.default_rng(n)
    a = np.asfortranarray(rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)).astype(cdt)
    t = _falloc(torch, shape, tcd)
    out = _falloc(torch, shape, tcd)
    for isign in (-1, 1):
        t.copy_(torch.from_numpy(a))
        d2d.c2c_1m(t, axis, isign, out=out)
        ref = orc.c2c_1m(a.astype(np.complex128), axis, isign)
        exact = np.fft.fft(a.astype(np.complex128), axis=axis) if isign == -1 else np.fft.ifft(a.astype(np.complex128), axis=axis) * n
        assert _relerr(out.cpu().numpy(), exact) < tol, (n, axis, isign)
        assert _relerr(out.cpu().numpy(), ref) < max(tol, 1e-10), (n, axis, isign)
    if axis != 1:
        r = np.asfortranarray(rng.uniform(-1, 1, shape)).astype(rdt)
        cs = list(shape)
        cs[axis] = n // 2 + 1
        tr, tc = _falloc(torch, shape, trd), _falloc(torch, tuple(cs), tcd)
        tr.copy_(torch.from_numpy(r))
        d2d.r2c_1m(tr, tc, axis)
        spec = np.fft.rfft(r.astype(np.float64), axis=axis)
        assert _relerr(tc.cpu().numpy(), spec) < tol, (n, axis, "r2c")
        assert _relerr(tc.cpu().numpy(), orc.r2c_1m(r.astype(np.float64), axis)) < max(tol, 1e-10), (n, axis, "r2c vs oracle")
        back = _falloc(torch, shape, trd)
        tc.copy_(torch.from_numpy(spec.astype(cdt)))
        d2d.c2r_1m(tc, back, axis)
        assert _relerr(back.cpu().numpy(), np.fft.irfft(spec, n=n, axis=axis) * n) < tol, (n, axis, "c2r")
    d2d.finalize()


@pytest.mark.parametrize("n,prec", [(1021, "f64"), (2042, "f64"), (4271, "f64"), (4801, "f64"), (9602, "f64"), (16411, "f64"),
                                    (1021, "f32"), (8537, "f32"), (17074, "f32")])
@pytest.mark.parametrize("axis", [0, 1, 2])
def test_large_prime_factors_chirp_z(n, prec, axis):
    """lengths with a prime factor beyond a few hundred -- including primes above the shared-memory kernel (fp64: 4265,
    fp32: 8532), which used to be rejected -- run as a circular convolution of power-of-two length (Bluestein): 1021 through
    the compiled 2048-point kernels, 4801 / 9602 / 16411 through the two-kernel split of 16384 / 32768.  The reference's
    generic backend takes any length (src/glassman.f90:29-67).  Pinned against the exact DFT (numpy's pocketfft); the
    Glassman oracle's recurrence twiddles drift to ~1e-11 at these primes (see tests/test_gpu_fft_any.py)."""
    p, d2d, torch = _ctx()
    cdt, rdt = (np.complex128, np.float64) if prec == "f64" else (np.complex64, np.float32)
    tcd, trd = (torch.complex128, torch.float64) if prec == "f64" else (torch.complex64, torch.float32)
    tol = 1e-12 if prec == "f64" else 1e-5
    shape = [3, 4, 2]
    shape[axis] = n
    shape = tuple(shape)
    rng = np.random.default_rng(n)
    a = np.asfortranarray(rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)).astype(cdt)
    t = _falloc(torch, shape, tcd)
    out = _falloc(torch, shape, tcd)
    for isign in (-1, 1):
        t.copy_(torch.from_numpy(a))
        d2d.c2c_1m(t, axis, isign, out=out)
        exact = np.fft.fft(a.astype(np.complex128), axis=axis) if isign == -1 else np.fft.ifft(a.astype(np.complex128), axis=axis) * n
        assert _relerr(out.cpu().numpy(), exact) < tol, (n, axis, isign)
        assert np.array_equal(t.cpu().numpy(), a), "input must be preserved when out != in"
        d2d.c2c_1m(t, axis, isign)  # in place
        assert _relerr(t.cpu().numpy(), exact) < tol, (n, axis, isign, "in place")
    if axis != 1:
        r = np.asfortranarray(rng.uniform(-1, 1, shape)).astype(rdt)
        cs = list(shape)
        cs[axis] = n // 2 + 1
        tr, tc = _falloc(torch, shape, trd), _falloc(torch, tuple(cs), tcd)
        tr.copy_(torch.from_numpy(r))
        d2d.r2c_1m(tr, tc, axis)
        spec = np.fft.rfft(r.astype(np.float64), axis=axis)
        assert _relerr(tc.cpu().numpy(), spec) < tol, (n, axis, "r2c")
        back = _falloc(torch, shape, trd)
        noisy = spec.copy()  # Im(bin 0) and (even n) Im(bin n/2) are ignored by c2r, like the reference's "real part of a c2c"
        idx = [slice(None)] * 3
        idx[axis] = 0
        noisy[tuple(idx)] += 0.5j
        if n % 2 == 0:
            idx[axis] = n // 2
            noisy[tuple(idx)] -= 0.25j
        tc.copy_(torch.from_numpy(np.asfortranarray(noisy).astype(cdt)))
        d2d.c2r_1m(tc, back, axis)
        assert _relerr(back.cpu().numpy(), np.fft.irfft(spec, n=n, axis=axis) * n) < tol, (n, axis, "c2r")
    d2d.finalize()


def test_fft_3d_with_a_large_prime_axis():
    """a 3-D r2c / c2r chain on a 2x2 grid whose y extent is a prime above the threshold (1021): the chirp-z path reads and
    writes through the multi-piece maps of the wire layouts like every other stage."""
    import oracle as orc2
    from util import run_ranks
    import torch
    p = pkg()
    shape, grid = (12, 1021, 10), (2, 2)
    rng = np.random.default_rng(5)
    g = np.asfortranarray(rng.uniform(-1, 1, shape))
    ins = orc2.scatter(g, grid, 0)
    exact = np.fft.fftn(g)[: shape[0] // 2 + 1]

    def body(rank, group):
        d2d = p.Decomp2d(*shape, grid[0], grid[1], rank=rank, nranks=4, group=group, device=0)
        eng = p.Decomp2dFFTEngine(d2d, p.PHYSICAL_IN_X)
        in_r, out_c = d2d.alloc_x(torch.float64, eng.ph), d2d.alloc_z(torch.complex128, eng.sp)
        in_r.copy_(torch.from_numpy(ins[rank]))
        eng.fft_3d(in_r, out_c)
        spec = out_c.cpu().numpy()
        back = d2d.alloc_x(torch.float64, eng.ph)
        eng.fft_3d(out_c, back)
        res = (spec, back.cpu().numpy(), tuple(eng.sp.zst), tuple(eng.sp.zen))
        eng.fin()
        d2d.finalize()
        return res

    res = run_ranks(4, body)
    smax = np.max(np.abs(exact))
    for r in range(4):
        spec, _, zst, zen = res[r]
        ref = exact[zst[0] - 1:zen[0], zst[1] - 1:zen[1], zst[2] - 1:zen[2]]
        assert np.max(np.abs(spec - ref)) / smax < 1e-12, r
    rt = orc2.gather([x[1] for x in res], shape, grid, 0) / np.prod(shape)
    assert np.max(np.abs(rt - g)) < 1e-13
