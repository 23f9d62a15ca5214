"""GPU parity of the 3-D transforms and the transposes (through the C ABI) against the oracle.

Multi-rank grids run as one thread per rank on one GPU (d2d_group transport: the exchange is a
device-to-device copy), which exercises exactly the piece maps / pack / unpack of the NCCL path.
The restated reference tests: examples/test2d/test2d.f90 (exact transposes),
examples/fft_physical_{x,z}/fft_{c2c,r2c}_{x,z}.f90 (round trip <= eps*50), fft_*_skip.f90.
"""
import numpy as np
import pytest

import oracle as orc
from util import pkg, run_ranks

pytestmark = pytest.mark.gpu

GRIDS = [(1, 1), (1, 2), (2, 1), (2, 2), (2, 4), (4, 2)]
TOL = {"f64": 1e-12, "f32": 1e-5}


def _np_dtypes(prec):
    return (np.float64, np.complex128) if prec == "f64" else (np.float32, np.complex64)


def _relerr(a, b):
    return np.max(np.abs(a - b)) / np.max(np.abs(b))


def _index_field(shape, dtype):
    nx, ny, nz = shape
    i, j, k = np.meshgrid(np.arange(1, nx + 1), np.arange(1, ny + 1), np.arange(1, nz + 1), indexing="ij")
    m = (i + (j - 1) * nx + (k - 1) * nx * ny).astype(np.float64)
    if np.dtype(dtype).kind == "c":
        return (m + 1j * (m - 1)).astype(dtype)
    return m.astype(dtype)


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.complex128, np.complex64])
@pytest.mark.parametrize("grid", GRIDS)
@pytest.mark.parametrize("shape", [(17, 13, 11), (32, 16, 64), (34, 26, 22)])
def test_test2d_transposes_bit_exact(shape, grid, dtype):
    """examples/test2d/test2d.f90:92-199 (+ timing2d_complex.f90 for complex): x->y->z->y->x, exact."""
    import torch
    p = pkg()
    g = _index_field(shape, dtype)
    want = [orc.scatter(g, grid, pen) for pen in range(3)]
    tdt = {np.float64: torch.float64, np.float32: torch.float32, np.complex128: torch.complex128, np.complex64: torch.complex64}[dtype]

    def body(rank, group):
        d2d = p.Decomp2d(*shape, grid[0], grid[1], rank=rank, nranks=grid[0] * grid[1], group=group, device=0)
        u1, u2, u3 = d2d.alloc_x(tdt), d2d.alloc_y(tdt), d2d.alloc_z(tdt)
        u1.copy_(torch.from_numpy(want[0][rank]))
        d2d.transpose_x_to_y(u1, u2)
        assert np.array_equal(u2.cpu().numpy(), want[1][rank]), "x->y"
        d2d.transpose_y_to_z(u2, u3)
        assert np.array_equal(u3.cpu().numpy(), want[2][rank]), "y->z"
        u2.zero_()
        d2d.transpose_z_to_y(u3, u2)
        assert np.array_equal(u2.cpu().numpy(), want[1][rank]), "z->y"
        u1.zero_()
        d2d.transpose_y_to_x(u2, u1)
        assert np.array_equal(u1.cpu().numpy(), want[0][rank]), "y->x"
        d2d.finalize()
        return True

    assert all(run_ranks(grid[0] * grid[1], body))


@pytest.mark.parametrize("dtype", [np.float64, np.complex64])
@pytest.mark.parametrize("grid", [(1, 2), (2, 2), (2, 4), (4, 2)])
@pytest.mark.parametrize("shape", [(17, 13, 11), (34, 26, 22), (32, 16, 64)])
def test_test2d_transposes_even_layout(shape, grid, dtype):
    """the same four transposes through the padded equal-count buffers of the reference's EVEN builds (d2d_ctx_set_even):
    ragged grids pad, the pencils are bit-identical"""
    import torch
    p = pkg()
    g = _index_field(shape, dtype)
    want = [orc.scatter(g, grid, pen) for pen in range(3)]
    tdt = {np.float64: torch.float64, np.complex64: torch.complex64}[dtype]

    def body(rank, group):
        d2d = p.Decomp2d(*shape, grid[0], grid[1], rank=rank, nranks=grid[0] * grid[1], group=group, device=0)
        d2d.set_even(True)
        u1, u2, u3 = d2d.alloc_x(tdt), d2d.alloc_y(tdt), d2d.alloc_z(tdt)
        u1.copy_(torch.from_numpy(want[0][rank]))
        d2d.transpose_x_to_y(u1, u2)
        assert np.array_equal(u2.cpu().numpy(), want[1][rank]), "x->y"
        d2d.transpose_y_to_z(u2, u3)
        assert np.array_equal(u3.cpu().numpy(), want[2][rank]), "y->z"
        u2.zero_()
        d2d.transpose_z_to_y(u3, u2)
        assert np.array_equal(u2.cpu().numpy(), want[1][rank]), "z->y"
        u1.zero_()
        d2d.transpose_y_to_x(u2, u1)
        assert np.array_equal(u1.cpu().numpy(), want[0][rank]), "y->x"
        d2d.finalize()
        return True

    assert all(run_ranks(grid[0] * grid[1], body))


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("fmt", [orc.PHYSICAL_IN_X, orc.PHYSICAL_IN_Z])
@pytest.mark.parametrize("grid", GRIDS)
@pytest.mark.parametrize("shape", [(32, 16, 64), (64, 64, 64), (16, 128, 8)])
def test_fft_3d_r2c_c2r(shape, grid, fmt, prec):
    """fft_r2c_x.f90 / fft_r2c_z.f90: forward spectrum vs oracle, c2r of the oracle spectrum vs oracle,
    round trip error per point <= eps*50; input of r2c is preserved."""
    import torch
    p = pkg()
    rdt, cdt = _np_dtypes(prec)
    trd, tcd = (torch.float64, torch.complex128) if prec == "f64" else (torch.float32, torch.complex64)
    rng = np.random.default_rng(7)
    g = np.asfortranarray(rng.uniform(-1, 1, shape)).astype(rdt)
    pin, pout = (0, 2) if fmt == orc.PHYSICAL_IN_X else (2, 0)
    ins = orc.scatter(g, grid, pin)
    ref_spec = orc.fft_3d_r2c_world(shape, grid, fmt, [a.astype(np.float64) for a in ins])
    ref_back = orc.fft_3d_c2r_world(shape, grid, fmt, ref_spec)
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
        in_c = alloc_out(tcd, eng.sp)
        in_c.copy_(torch.from_numpy(ref_spec[rank].astype(cdt)))
        back = alloc_in(trd, eng.ph)
        eng.fft_3d(in_c, back)
        assert np.array_equal(in_c.cpu().numpy(), ref_spec[rank].astype(cdt)), "c2r (not inplace) must preserve its input"
        rt = alloc_in(trd, eng.ph)
        eng.fft_3d(out_c, rt)
        res = (spec, back.cpu().numpy(), rt.cpu().numpy())
        eng.fin()
        d2d.finalize()
        return res

    res = run_ranks(nranks, body)
    smax = max(np.max(np.abs(s)) for s in ref_spec)
    bmax = max(np.max(np.abs(s)) for s in ref_back)
    for r in range(nranks):
        assert np.max(np.abs(res[r][0] - ref_spec[r])) / smax < TOL[prec], ("spectrum", r)
        assert np.max(np.abs(res[r][1] - ref_back[r])) / bmax < TOL[prec], ("c2r", r)
    rt = orc.gather([x[2] for x in res], shape, grid, pin).astype(np.float64) / np.prod(shape)
    assert np.sum(np.abs(rt - g)) / np.prod(shape) < np.finfo(rdt).eps * 50


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("fmt", [orc.PHYSICAL_IN_X, orc.PHYSICAL_IN_Z])
@pytest.mark.parametrize("grid", [(1, 1), (1, 2), (2, 2), (2, 4)])
@pytest.mark.parametrize("inplace", [False, True])
def test_fft_3d_c2c(grid, fmt, prec, inplace):
    """fft_c2c_x.f90 / fft_c2c_z.f90:64-154 with the reference's (1+i) ramp field and a random field."""
    import torch
    p = pkg()
    shape = (32, 64, 16)
    rdt, cdt = _np_dtypes(prec)
    tcd = torch.complex128 if prec == "f64" else torch.complex64
    rng = np.random.default_rng(11)
    g = np.asfortranarray(rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)).astype(cdt)
    pin, pout = (0, 2) if fmt == orc.PHYSICAL_IN_X else (2, 0)
    ins = orc.scatter(g, grid, pin)
    ref_spec = orc.fft_3d_c2c_world(shape, grid, fmt, orc.FORWARD, [a.astype(np.complex128) for a in ins])
    nranks = grid[0] * grid[1]

    def body(rank, group):
        d2d = p.Decomp2d(*shape, grid[0], grid[1], rank=rank, nranks=nranks, group=group, device=0)
        eng = p.Decomp2dFFTEngine(d2d, fmt, dtype=torch.float64 if prec == "f64" else torch.float32, opt_inplace=inplace)
        a_in = (d2d.alloc_x if pin == 0 else d2d.alloc_z)(tcd)
        a_out = (d2d.alloc_x if pout == 0 else d2d.alloc_z)(tcd)
        a_in.copy_(torch.from_numpy(ins[rank]))
        eng.fft_3d(a_in, a_out, p.DECOMP_2D_FFT_FORWARD)
        if not inplace:
            assert np.array_equal(a_in.cpu().numpy(), ins[rank]), "c2c (not inplace) must preserve its input"
        spec = a_out.cpu().numpy()
        a_back = (d2d.alloc_x if pin == 0 else d2d.alloc_z)(tcd)
        eng.fft_3d(a_out, a_back, p.DECOMP_2D_FFT_BACKWARD)
        res = (spec, a_back.cpu().numpy())
        eng.fin()
        d2d.finalize()
        return res

    res = run_ranks(nranks, body)
    smax = max(np.max(np.abs(s)) for s in ref_spec)
    for r in range(nranks):
        assert np.max(np.abs(res[r][0] - ref_spec[r])) / smax < TOL[prec]
    rt = orc.gather([x[1] for x in res], shape, grid, pin) / np.prod(shape)
    assert np.sum(np.abs(rt - g)) / np.prod(shape) < np.finfo(rdt).eps * 50


@pytest.mark.parametrize("skip", [(True, False, True), (False, True, False)])
@pytest.mark.parametrize("grid", [(1, 1), (2, 2)])
def test_fft_c2c_skip_flags(skip, grid):
    """fft_c2c_x_skip.f90: opt_skip_XYZ_c2c."""
    import torch
    p = pkg()
    shape = (16, 32, 8)
    rng = np.random.default_rng(3)
    g = np.asfortranarray(rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape))
    ins = orc.scatter(g, grid, 0)
    ref = orc.fft_3d_c2c_world(shape, grid, orc.PHYSICAL_IN_X, orc.FORWARD, ins, skip=skip)
    nranks = grid[0] * grid[1]

    def body(rank, group):
        d2d = p.Decomp2d(*shape, grid[0], grid[1], rank=rank, nranks=nranks, group=group, device=0)
        eng = p.Decomp2dFFTEngine(d2d, p.PHYSICAL_IN_X, opt_skip_XYZ_c2c=skip)
        a_in, a_out = d2d.alloc_x(torch.complex128), d2d.alloc_z(torch.complex128)
        a_in.copy_(torch.from_numpy(ins[rank]))
        eng.fft_3d(a_in, a_out, p.DECOMP_2D_FFT_FORWARD)
        out = a_out.cpu().numpy()
        eng.fin()
        d2d.finalize()
        return out

    res = run_ranks(nranks, body)
    for r in range(nranks):
        assert _relerr(res[r], ref[r]) < 1e-12


def test_fft_grid_differs_from_init_grid():
    """examples/fft_physical_x/fft_grid_x.f90: library initialised with (nx+1,ny+1,nz+1), FFT grid (nx,ny,nz)."""
    import torch
    p = pkg()
    shape, grid = (32, 16, 64), (2, 2)
    rng = np.random.default_rng(5)
    g = np.asfortranarray(rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape))
    ins = orc.scatter(g, grid, 0)
    ref = orc.fft_3d_c2c_world(shape, grid, orc.PHYSICAL_IN_X, orc.FORWARD, ins)

    def body(rank, group):
        d2d = p.Decomp2d(shape[0] + 1, shape[1] + 1, shape[2] + 1, grid[0], grid[1], rank=rank, nranks=4, group=group, device=0)
        eng = p.Decomp2dFFTEngine(d2d, p.PHYSICAL_IN_X, *shape)
        a_in, a_out = d2d.alloc_x(torch.complex128, eng.ph), d2d.alloc_z(torch.complex128, eng.ph)
        a_in.copy_(torch.from_numpy(ins[rank]))
        eng.fft_3d(a_in, a_out, p.DECOMP_2D_FFT_FORWARD)
        out = a_out.cpu().numpy()
        eng.fin()
        d2d.finalize()
        return out

    res = run_ranks(4, body)
    for r in range(4):
        assert _relerr(res[r], ref[r]) < 1e-12


def test_module_level_api_and_host_arrays():
    """The reference-named module-level calls + the host-array entry points (e2e path of bench.py)."""
    import torch
    p = pkg()
    shape = (64, 32, 128)
    d2d = p.decomp_2d_init(*shape, 1, 1)
    eng = p.decomp_2d_fft_init(p.PHYSICAL_IN_Z)
    assert p.decomp_2d_fft_get_size()[2] == (64, 32, 65)
    rng = np.random.default_rng(0)
    g = np.asfortranarray(rng.uniform(-1, 1, shape))
    ref = orc.fft_3d_r2c_world(shape, (1, 1), orc.PHYSICAL_IN_Z, [g])[0]
    h_in = torch.from_numpy(g.copy(order="F")).pin_memory() if False else None
    in_h = np.asfortranarray(g)
    out_h = np.zeros(ref.shape, dtype=np.complex128, order="F")
    eng.fft_3d_r2c_host(in_h.ctypes.data, out_h.ctypes.data)
    assert _relerr(out_h, ref) < 1e-12
    back_h = np.zeros(shape, dtype=np.float64, order="F")
    eng.fft_3d_c2r_host(out_h.ctypes.data, back_h.ctypes.data)
    assert np.max(np.abs(back_h / np.prod(shape) - g)) < 1e-13
    p.decomp_2d_finalize()


def test_rank_failure_aborts_the_group_instead_of_hanging():
    """One rank fails before an exchange: the ranks waiting for it get an error (d2d_group_abort), like
    decomp_2d_abort -> MPI_ABORT takes the whole job down in the reference (src/decomp_2d_mpi.f90:166-191)."""
    import time
    import torch
    p = pkg()
    shape, grid = (16, 16, 16), (2, 2)

    def body(rank, group):
        d2d = p.Decomp2d(*shape, grid[0], grid[1], rank=rank, nranks=4, group=group, device=0)
        if rank == 3:
            raise RuntimeError("injected failure on rank 3")
        u1, u2 = d2d.alloc_x(torch.float64), d2d.alloc_y(torch.float64)
        d2d.transpose_x_to_y(u1, u2)  # ranks 1 (same column as 3) waits for rank 3 in the exchange
        return True

    t0 = time.monotonic()
    with pytest.raises(RuntimeError, match="injected failure"):
        run_ranks(4, body, timeout=60)
    assert time.monotonic() - t0 < 30


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("fmt", [orc.PHYSICAL_IN_X, orc.PHYSICAL_IN_Z])
@pytest.mark.parametrize("grid", [(1, 1), (2, 2)])
@pytest.mark.parametrize("shape", [(34, 8, 256), (66, 6, 512), (256, 10, 36), (30, 6, 256)])
def test_fft_3d_r2c_odd_pitch_pencils_on_tma_kernels(shape, grid, fmt, prec):
    """Pencils whose row pitch is odd (nx/2+1 bins, 18- or 34-wide spectral pencils, 30-wide real ones) next to transform
    lengths >= 256, i.e. on the TMA-staged kernels: shifted tile columns, ragged first/last tiles, merged r2c landing with
    partial groups.  Forward spectrum against the oracle and round trip."""
    import torch
    p = pkg()
    rdt = np.float64 if prec == "f64" else np.float32
    trd, tcd = (torch.float64, torch.complex128) if prec == "f64" else (torch.float32, torch.complex64)
    rng = np.random.default_rng(13)
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


@pytest.mark.parametrize("fmt", [orc.PHYSICAL_IN_X, orc.PHYSICAL_IN_Z])
@pytest.mark.parametrize("grid", [(1, 2), (2, 1), (2, 2), (2, 4)])
@pytest.mark.parametrize("shape", [(64, 48, 32), (256, 40, 36), (34, 26, 22)])
def test_overlapped_chain_matches_oracle(shape, grid, fmt, monkeypatch):
    """D2D_OVERLAP=3: every multi-rank link is cut into three chunks along its free axis, producer chunks / exchanges /
    consumer chunks pipelined on two streams (run_chain_overlap).  Same results as the sequential chain: r2c spectrum and
    c2c spectrum against the oracle, round trip, repeated calls (buffer rotation across calls)."""
    import torch
    monkeypatch.setenv("D2D_OVERLAP", "3")
    p = pkg()
    rng = np.random.default_rng(17)
    g = np.asfortranarray(rng.uniform(-1, 1, shape))
    gc = np.asfortranarray(rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape))
    pin, pout = (0, 2) if fmt == orc.PHYSICAL_IN_X else (2, 0)
    ins, cins = orc.scatter(g, grid, pin), orc.scatter(gc, grid, pin)
    ref_spec = orc.fft_3d_r2c_world(shape, grid, fmt, ins)
    ref_c = orc.fft_3d_c2c_world(shape, grid, fmt, orc.FORWARD, cins)
    nranks = grid[0] * grid[1]

    def body(rank, group):
        d2d = p.Decomp2d(*shape, grid[0], grid[1], rank=rank, nranks=nranks, group=group, device=0)
        eng = p.Decomp2dFFTEngine(d2d, fmt)
        alloc_in = d2d.alloc_x if pin == 0 else d2d.alloc_z
        alloc_out = d2d.alloc_z if pin == 0 else d2d.alloc_x
        in_r, out_c = alloc_in(torch.float64, eng.ph), alloc_out(torch.complex128, eng.sp)
        in_r.copy_(torch.from_numpy(ins[rank]))
        for _ in range(2):
            eng.fft_3d(in_r, out_c)
        spec = out_c.cpu().numpy()
        rt = alloc_in(torch.float64, eng.ph)
        eng.fft_3d(out_c, rt)
        c_in, c_out = alloc_in(torch.complex128, eng.ph), alloc_out(torch.complex128, eng.ph)
        c_in.copy_(torch.from_numpy(cins[rank]))
        eng.fft_3d(c_in, c_out, p.DECOMP_2D_FFT_FORWARD)
        res = (spec, rt.cpu().numpy(), c_out.cpu().numpy())
        eng.fin()
        d2d.finalize()
        return res

    res = run_ranks(nranks, body)
    smax = max(np.max(np.abs(s)) for s in ref_spec if s.size)
    cmax = max(np.max(np.abs(s)) for s in ref_c if s.size)
    for r in range(nranks):
        if ref_spec[r].size:
            assert np.max(np.abs(res[r][0] - ref_spec[r])) / smax < 1e-12, ("r2c", r)
        if ref_c[r].size:
            assert np.max(np.abs(res[r][2] - ref_c[r])) / cmax < 1e-12, ("c2c", r)
    rt = orc.gather([x[1] for x in res], shape, grid, pin) / np.prod(shape)
    assert np.max(np.abs(rt - g)) < 1e-13


@pytest.mark.parametrize("grid", [(8, 1), (1, 8), (8, 2)])
def test_process_grid_sides_at_the_cap(grid):
    """the widest process-grid side the library takes (8 ranks per communicator = kMaxPieces pieces per line): transposes
    bit-exact and r2c / c2r against the oracle on a grid whose pencils are ragged (34 / 8, 18 / 8 spectral planes)."""
    import torch
    p = pkg()
    shape = (34, 40, 24)
    nranks = grid[0] * grid[1]
    g = _index_field(shape, np.float64)
    want = [orc.scatter(g, grid, pen) for pen in range(3)]
    rng = np.random.default_rng(8)
    f = np.asfortranarray(rng.uniform(-1, 1, shape))
    ins = orc.scatter(f, grid, 0)
    ref = orc.fft_3d_r2c_world(shape, grid, orc.PHYSICAL_IN_X, ins)

    def body(rank, group):
        d2d = p.Decomp2d(*shape, grid[0], grid[1], rank=rank, nranks=nranks, group=group, device=0)
        u1, u2, u3 = d2d.alloc_x(torch.float64), d2d.alloc_y(torch.float64), d2d.alloc_z(torch.float64)
        u1.copy_(torch.from_numpy(want[0][rank]))
        d2d.transpose_x_to_y(u1, u2)
        d2d.transpose_y_to_z(u2, u3)
        ok = np.array_equal(u2.cpu().numpy(), want[1][rank]) and np.array_equal(u3.cpu().numpy(), want[2][rank])
        eng = p.Decomp2dFFTEngine(d2d, p.PHYSICAL_IN_X)
        in_r, out_c = d2d.alloc_x(torch.float64, eng.ph), d2d.alloc_z(torch.complex128, eng.sp)
        in_r.copy_(torch.from_numpy(ins[rank]))
        eng.fft_3d(in_r, out_c)
        spec = out_c.cpu().numpy()
        back = d2d.alloc_x(torch.float64, eng.ph)
        eng.fft_3d(out_c, back)
        res = (ok, spec, back.cpu().numpy())
        eng.fin()
        d2d.finalize()
        return res

    res = run_ranks(nranks, body)
    assert all(r[0] for r in res), "transposes"
    smax = max(np.max(np.abs(s)) for s in ref if s.size)
    for r in range(nranks):
        if ref[r].size:
            assert np.max(np.abs(res[r][1] - ref[r])) / smax < 1e-12, r
    rt = orc.gather([x[2] for x in res], shape, grid, 0) / np.prod(shape)
    assert np.max(np.abs(rt - f)) < 1e-13


def test_process_grid_side_beyond_the_cap_is_rejected():
    p = pkg()
    grp = p.Group(9)
    with pytest.raises(p.Decomp2dError, match="larger than 8"):
        p.Decomp2d(64, 64, 64, 9, 1, rank=0, nranks=9, group=grp, device=0)
    grp.destroy()
