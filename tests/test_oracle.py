"""CPU tests of the oracle (oracle/d2d_oracle.c) against the reference's own fixtures.

The reference's tests generate their fixtures in code; each test below restates one of them and
cites it.  Forward spectra are not pinned by any reference test ("parity unpinned" there); they are
pinned here against numpy/pocketfft and the analytic DFT of the reference's ramp field.
"""
import numpy as np
import pytest

import oracle as orc

GRIDS = [(1, 1), (1, 2), (2, 1), (2, 2), (2, 4), (3, 2), (4, 2)]


def index_field(shape, dtype=np.float64):
    """examples/test2d/test2d.f90:74-90: u(i,j,k) = i + (j-1) nx + (k-1) nx ny (1-based)."""
    nx, ny, nz = shape
    i, j, k = np.meshgrid(np.arange(1, nx + 1), np.arange(1, ny + 1), np.arange(1, nz + 1), indexing="ij")
    m = (i + (j - 1) * nx + (k - 1) * nx * ny).astype(np.float64)
    if np.dtype(dtype).kind == "c":
        return (m + 1j * (m - 1)).astype(dtype)  # examples/test2d/timing2d_complex.f90:89-105
    return m.astype(dtype)


def ramp_field(shape, dtype=np.float64):
    """examples/fft_physical_x/fft_r2c_x.f90:66-76: in(i,j,k) = (i/nx)(j/ny)(k/nz), 1-based."""
    nx, ny, nz = shape
    i, j, k = np.meshgrid(np.arange(1, nx + 1), np.arange(1, ny + 1), np.arange(1, nz + 1), indexing="ij")
    return ((i / nx) * (j / ny) * (k / nz)).astype(dtype)


@pytest.mark.parametrize("grid", GRIDS + [(5, 3), (8, 1), (1, 8)])
@pytest.mark.parametrize("shape", [(17, 13, 11), (64, 64, 64), (33, 64, 9), (8, 8, 8)])
def test_init_test_size_conservation(shape, grid):
    """examples/init_test/init_test.f90:82-112."""
    nx, ny, nz = shape
    if min(nx, ny) < grid[0] or min(ny, nz) < grid[1]:
        pytest.skip("reference aborts for such grids (decomp_2d_init_fin.f90:43-45)")
    tot = [0, 0, 0]
    for r in range(grid[0] * grid[1]):
        d = orc.Decomp(nx, ny, nz, grid[0], grid[1], r)
        for p in range(3):
            tot[p] += int(np.prod(d.sz(p)))
            assert all(e - s + 1 == z for s, e, z in zip(d.st(p), (d.xen, d.yen, d.zen)[p], d.sz(p)))
    assert tot == [nx * ny * nz] * 3


def test_distribute_extras_go_last():
    """decomp_2d.f90:1070-1105, comment at :1108-1110: 17 over 4 -> (4,4,4,5)."""
    assert orc.distribute(17, 4)[2] == [4, 4, 4, 5]
    assert orc.distribute(513, 4)[2] == [128, 128, 128, 129]
    assert orc.distribute(257, 2)[2] == [128, 129]
    assert orc.distribute(16, 4) == ([1, 5, 9, 13], [4, 8, 12, 16], [4, 4, 4, 4])


def test_best_2d_grid():
    """decomp_2d_init_fin.f90:270-300 + factor.f90: col = factors(nfact/2+1)."""
    assert orc.best_2d_grid(8) == (2, 4)
    assert orc.best_2d_grid(4) == (2, 2)
    assert orc.best_2d_grid(2) == (1, 2)
    assert orc.best_2d_grid(1) == (1, 1)
    assert orc.best_2d_grid(12) == (3, 4)
    assert orc.best_2d_grid(16) == (4, 4)


@pytest.mark.parametrize("dtype", [np.float64, np.complex128, np.float32, np.complex64])
@pytest.mark.parametrize("grid", GRIDS)
@pytest.mark.parametrize("shape", [(17, 13, 11), (16, 32, 8), (34, 26, 22)])
def test_test2d_transposes_exact(shape, grid, dtype):
    """examples/test2d/test2d.f90:92-199: after EACH of x->y, y->z, z->y, y->x the field equals
    the index field of the destination pencil, exactly."""
    g = index_field(shape, dtype)
    u1 = orc.scatter(g, grid, 0)
    u2 = orc.transpose_world(orc.X_TO_Y, shape, grid, u1)
    for a, b in zip(u2, orc.scatter(g, grid, 1)):
        assert np.array_equal(a, b)
    u3 = orc.transpose_world(orc.Y_TO_Z, shape, grid, u2)
    for a, b in zip(u3, orc.scatter(g, grid, 2)):
        assert np.array_equal(a, b)
    u2b = orc.transpose_world(orc.Z_TO_Y, shape, grid, u3)
    for a, b in zip(u2b, orc.scatter(g, grid, 1)):
        assert np.array_equal(a, b)
    u1b = orc.transpose_world(orc.Y_TO_X, shape, grid, u2b)
    for a, b in zip(u1b, u1):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 8, 11, 13, 16, 17, 34, 51, 64, 100, 128, 256, 1000, 1024, 1025,
                               6, 12, 24, 48, 96, 192, 384, 768, 1536, 3072, 10, 20, 40, 80, 160, 320, 640, 1280, 510])  # + the compiled 3*2^k / 5*2^k lengths
@pytest.mark.parametrize("isign", [-1, 1])
def test_spcfft_vs_pocketfft(n, isign):
    """SPCFFT (glassman.f90:29-108) is an unnormalised DFT with exp(isign 2 pi i jk/n)."""
    rng = np.random.default_rng(n)
    u = rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)
    ref = np.fft.fft(u) if isign == -1 else np.fft.ifft(u) * n
    got = orc.spcfft(u, isign)
    assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < 2e-12
    got32 = orc.spcfft(u.astype(np.complex64), isign)
    assert got32.dtype == np.complex64
    assert np.max(np.abs(got32 - ref)) / np.max(np.abs(ref)) < 5e-4


def test_spcfft_analytic_ramp():
    """Known answer for the reference's example field (examples/fft_physical_x/fft_r2c_x.f90:66-76
    is a product of ramps u[m] = (m+1)/n, m=0..n-1): DFT[0] = (n+1)/2, DFT[k] = -1/(1 - w^k),
    w = exp(-2 pi i/n)  (geometric-series identity)."""
    for n in (11, 64, 1024):
        u = (np.arange(1, n + 1) / n).astype(np.complex128)
        got = orc.spcfft(u, -1)
        k = np.arange(1, n)
        w = np.exp(-2j * np.pi * k / n)
        ana = np.concatenate([[(n + 1) / 2], -1.0 / (1.0 - w)])
        assert np.max(np.abs(got - ana)) / np.max(np.abs(ana)) < 1e-11


@pytest.mark.parametrize("fmt", [orc.PHYSICAL_IN_X, orc.PHYSICAL_IN_Z])
@pytest.mark.parametrize("grid", GRIDS)
@pytest.mark.parametrize("shape", [(17, 13, 11), (16, 32, 8)])
def test_fft_3d_c2c_forward_spectrum_and_roundtrip(shape, grid, fmt):
    """fft_c2c_x.f90 / fft_c2c_z.f90:64-154 (round trip <= eps*50*ntest) + forward spectrum vs pocketfft."""
    rng = np.random.default_rng(1)
    g = (rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)).astype(np.complex128)
    pin, pout = (0, 2) if fmt == orc.PHYSICAL_IN_X else (2, 0)
    ins = orc.scatter(g, grid, pin)
    outs = orc.fft_3d_c2c_world(shape, grid, fmt, orc.FORWARD, ins)
    spec = orc.gather(outs, shape, grid, pout)
    ref = np.fft.fftn(g)
    assert np.max(np.abs(spec - ref)) / np.max(np.abs(ref)) < 1e-12
    back = orc.fft_3d_c2c_world(shape, grid, fmt, orc.BACKWARD, outs)
    rt = orc.gather(back, shape, grid, pin) / np.prod(shape)
    err = np.sum(np.abs(rt - g)) / np.prod(shape)
    assert err < np.finfo(np.float64).eps * 50


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("fmt", [orc.PHYSICAL_IN_X, orc.PHYSICAL_IN_Z])
@pytest.mark.parametrize("grid", GRIDS)
@pytest.mark.parametrize("shape", [(17, 13, 11), (16, 32, 8), (12, 10, 14)])
def test_fft_3d_r2c_c2r(shape, grid, fmt, dtype):
    """fft_r2c_x.f90:66-152 / fft_r2c_z.f90: ramp field, round-trip L1 error/point <= eps*50*ntest;
    plus forward half-spectrum vs pocketfft rfftn (axis 0 halved for X, axis 2 for Z)."""
    g = ramp_field(shape, dtype)
    pin, pout = (0, 2) if fmt == orc.PHYSICAL_IN_X else (2, 0)
    sps = orc.sp_shape(shape, fmt)
    ins = orc.scatter(g, grid, pin)
    outs = orc.fft_3d_r2c_world(shape, grid, fmt, ins)
    spec = orc.gather(outs, sps, grid, pout)
    full = np.fft.fftn(g.astype(np.float64))
    ref = full[: sps[0], :, :] if fmt == orc.PHYSICAL_IN_X else full[:, :, : sps[2]]
    tol = 1e-12 if dtype == np.float64 else 2e-5
    assert np.max(np.abs(spec - ref)) / np.max(np.abs(ref)) < tol
    back = orc.fft_3d_c2r_world(shape, grid, fmt, outs)
    rt = orc.gather(back, shape, grid, pin) / np.prod(shape)
    err = np.sum(np.abs(rt.astype(np.float64) - g)) / np.prod(shape)
    assert err < np.finfo(dtype).eps * 50


@pytest.mark.parametrize("skip", [(True, False, True), (False, True, False)])
def test_fft_c2c_skip_flags(skip):
    """fft_c2c_x_skip.f90: opt_skip_XYZ_c2c -> only the non-skipped axes are transformed."""
    shape, grid = (12, 10, 8), (2, 2)
    rng = np.random.default_rng(3)
    g = (rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)).astype(np.complex128)
    outs = orc.fft_3d_c2c_world(shape, grid, orc.PHYSICAL_IN_X, orc.FORWARD, orc.scatter(g, grid, 0), skip=skip)
    spec = orc.gather(outs, shape, grid, 2)
    axes = [a for a in range(3) if not skip[a]]
    ref = np.fft.fftn(g, axes=axes)
    assert np.max(np.abs(spec - ref)) / np.max(np.abs(ref)) < 1e-12


def test_c2r_uses_imag_of_dc_and_nyquist_like_reference():
    """fft_generic.f90:320-337: bins 0 and n/2 keep their imaginary parts in the c2c(+1); the REAL
    part of the result is taken, so those imaginary parts cannot reach the output."""
    n = 16
    rng = np.random.default_rng(5)
    a = (rng.uniform(-1, 1, (n // 2 + 1, 2, 2)) + 1j * rng.uniform(-1, 1, (n // 2 + 1, 2, 2))).astype(np.complex128)
    got = orc.c2r_1m(a, n, 0)
    a2 = a.copy()
    a2[0].imag = 0
    a2[n // 2].imag = 0
    ref = np.fft.irfft(a2, n=n, axis=0) * n
    assert np.max(np.abs(got - ref)) < 1e-12


@pytest.mark.parametrize("pencil", [0, 1, 2])
@pytest.mark.parametrize("grid", [(1, 1), (2, 2), (2, 3)])
def test_halo_oracle_periodic_closed_form(pencil, grid):
    """with every axis periodic each ghost cell of update_halo is the global field at the wrapped index (the property
    examples/halo_test relies on); without periodicity the ghost layers beyond the domain stay untouched"""
    shape, level = (7, 6, 9), 2
    g = np.arange(1, np.prod(shape) + 1, dtype=np.float64).reshape(shape, order="F")
    outs = orc.update_halo_world(g, grid, pencil, level, (True, True, True))
    axes = [a for a in range(3) if a != pencil]
    for r, o in enumerate(outs):
        d = orc.Decomp(*shape, grid[0], grid[1], r)
        st = d.st(pencil)
        idx = [np.arange(o.shape[a]) + st[a] - (level if a in axes else 0) for a in range(3)]
        want = g[np.ix_(idx[0] % shape[0], idx[1] % shape[1], idx[2] % shape[2])]
        assert np.array_equal(o, want), (pencil, grid, r)
    outs = orc.update_halo_world(g, grid, pencil, level)
    for r, o in enumerate(outs):
        d = orc.Decomp(*shape, grid[0], grid[1], r)
        st = d.st(pencil)
        idx = [np.arange(o.shape[a]) + st[a] - (level if a in axes else 0) for a in range(3)]
        inside = np.ix_(*[(i >= 0) & (i < shape[a]) for a, i in enumerate(idx)])
        want = np.zeros_like(o)
        want[inside] = g[np.ix_(*[i[(i >= 0) & (i < shape[a])] for a, i in enumerate(idx)])]
        assert np.array_equal(o, want), (pencil, grid, r)
