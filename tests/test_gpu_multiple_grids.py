"""examples/fft_multiple_grids restated (fft_multiple_grids.f90:37-66, fft_multiple_grids_utilities.f90:15-193 ntest_c2c,
196-395 ntest_r2c, fft_multiple_grids_reference.f90): several live FFT engines on grids that differ from the grid of
decomp_2d_init, switched with decomp_2d_fft_use_grid, plus a local engine object;

    c2c field (r, -3 r), r = (i/nx)(j/ny)(k/nz) with GLOBAL indices; forward, the input array is ZEROED, backward into it,
    rescale; after nt rounds  sqrt(sum |err|^2) / (nx ny nz) <= epsilon * 5 * nt    (utilities.f90:157-191)
    r2c the same with the real field r                                               (utilities.f90:330-393)

and the engine API around it (src/fft_common.f90:368-555): set_ngrid / get_ngrid / use_grid / get_engine / get_format /
get_inplace[_r2c/_c2r] with their error behaviour."""
import numpy as np
import pytest

from util import pkg, run_ranks

pytestmark = pytest.mark.gpu


def _ramp(torch, st, sz, n, dev):
    ax = [torch.arange(st[d], st[d] + sz[d], device=dev, dtype=torch.float64) / n[d] for d in range(3)]
    return ax[0][:, None, None] * ax[1][None, :, None] * ax[2][None, None, :]


def _ntest_c2c(p, torch, d2d, eng, nt):
    ph = eng.ph
    n = (eng.nx_fft, eng.ny_fft, eng.nz_fft)
    fx = eng.format == p.PHYSICAL_IN_X
    a_in, a_out = (d2d.alloc_x, d2d.alloc_z) if fx else (d2d.alloc_z, d2d.alloc_x)
    inp = a_in(eng.complex_dtype, ph, opt_global=True)
    out = a_out(eng.complex_dtype, ph, opt_global=True)
    st, sz = (ph.xst, ph.xsz) if fx else (ph.zst, ph.zsz)
    assert inp.lbound == tuple(st)
    r = _ramp(torch, st, sz, n, inp.device)
    inp.copy_(torch.complex(r, -3 * r))
    for _ in range(nt + 1):  # the example runs one iteration outside its timed loop
        eng.fft_3d(inp, out, p.DECOMP_2D_FFT_FORWARD)
        inp.zero_()
        eng.fft_3d(out, inp, p.DECOMP_2D_FFT_BACKWARD)
        inp.div_(float(n[0]) * n[1] * n[2])
    return float(((inp.real - r) ** 2 + (inp.imag + 3 * r) ** 2).sum().item())


def _ntest_r2c(p, torch, d2d, eng, nt):
    ph, sp = eng.ph, eng.sp
    n = (eng.nx_fft, eng.ny_fft, eng.nz_fft)
    fx = eng.format == p.PHYSICAL_IN_X
    a_in, a_out = (d2d.alloc_x, d2d.alloc_z) if fx else (d2d.alloc_z, d2d.alloc_x)
    inp = a_in(eng.real_dtype, ph, opt_global=True)
    out = a_out(eng.complex_dtype, sp, opt_global=True)
    st, sz = (ph.xst, ph.xsz) if fx else (ph.zst, ph.zsz)
    r = _ramp(torch, st, sz, n, inp.device)
    inp.copy_(r)
    for _ in range(nt + 1):
        eng.fft_3d(inp, out)
        inp.zero_()
        eng.fft_3d(out, inp)
        inp.div_(float(n[0]) * n[1] * n[2])
    return float(((inp - r) ** 2).sum().item())


@pytest.mark.parametrize("base", [(17, 13, 11), (32, 16, 64)])
@pytest.mark.parametrize("grid", [(1, 1), (2, 2)])
def test_fft_multiple_grids(base, grid):
    import torch
    p = pkg()
    nx, ny, nz = base
    nt = 3
    nranks = grid[0] * grid[1]

    def body(rank, group):
        # fft_multiple_grids.f90:43: the decomposition library is initialised on a grid no FFT engine uses
        d2d = p.decomp_2d_init(nx + 1, ny + 1, nz + 1, grid[0], grid[1], rank=rank, nranks=nranks, group=group, device=0)
        p.decomp_2d_fft_set_ngrid(3)
        assert p.decomp_2d_fft_get_ngrid() == 3
        p.decomp_2d_fft_init(p.PHYSICAL_IN_X, nx, ny, nz, 1)
        p.decomp_2d_fft_init(p.PHYSICAL_IN_Z, nx, ny, nz, 2)
        p.decomp_2d_fft_init(p.PHYSICAL_IN_X, nx, ny + 2, nz + 16, 3)
        local = p.Decomp2dFFTEngine(d2d, p.PHYSICAL_IN_Z, nx + 16, ny + 2, nz)  # local_engine%init (:54)
        sums = []
        for igrid in (1, 2, 3):
            eng = p.decomp_2d_fft_use_grid(igrid)
            assert eng is p.decomp_2d_fft_get_engine(igrid)
            assert p.decomp_2d_fft_get_format() == (p.PHYSICAL_IN_Z if igrid == 2 else p.PHYSICAL_IN_X)
            assert p.decomp_2d_fft_get_inplace() is False and p.decomp_2d_fft_get_inplace_r2c() is False
            assert p.decomp_2d_fft_get_inplace_c2r() is False
            assert tuple(p.decomp_2d_fft_get_ph().xsz)[0] == eng.nx_fft
            sums.append((eng, _ntest_c2c(p, torch, d2d, eng, nt), _ntest_r2c(p, torch, d2d, eng, nt)))
        sums.append((local, _ntest_c2c(p, torch, d2d, local, nt), _ntest_r2c(p, torch, d2d, local, nt)))
        out = [((e.nx_fft, e.ny_fft, e.nz_fft), c, r) for e, c, r in sums]
        local.fin()
        p.decomp_2d_finalize()
        return out

    res = run_ranks(nranks, body) if nranks > 1 else [body(0, None)]
    eps = np.finfo(np.float64).eps
    for g in range(4):
        n = res[0][g][0]
        npts = float(n[0]) * n[1] * n[2]
        err_c = np.sqrt(sum(r[g][1] for r in res)) / npts  # MPI_ALLREDUCE(SUM) of the squared errors, then sqrt / N
        err_r = np.sqrt(sum(r[g][2] for r in res)) / npts
        assert err_c <= eps * 5 * nt, (n, err_c)
        assert err_r <= eps * 5 * nt, (n, err_r)


def test_engine_api_errors():
    """src/fft_common.f90:93-95 (igrid outside 1..n_grid), :383-385 (n_grid < 1), :425-437 (use_grid), :177-182 (in-place
    r2c / c2r are refused by this backend), precision and aliasing checks of the mirror"""
    import torch
    p = pkg()
    d2d = p.decomp_2d_init(16, 16, 16, 1, 1)
    with pytest.raises(p.Decomp2dError, match="Invalid value for n_grid"):
        p.decomp_2d_fft_set_ngrid(0)
    p.decomp_2d_fft_set_ngrid(2)
    with pytest.raises(p.Decomp2dError, match="Invalid value for igrid"):
        p.decomp_2d_fft_init(p.PHYSICAL_IN_X, 16, 16, 16, 3)
    with pytest.raises(p.Decomp2dError, match="not ready"):
        p.decomp_2d_fft_use_grid(2)
    with pytest.raises(p.Decomp2dError, match="In-place r2c"):
        p.decomp_2d_fft_init(p.PHYSICAL_IN_X, 16, 16, 16, 1, opt_inplace_r2c=True)
    eng = p.decomp_2d_fft_init(p.PHYSICAL_IN_X, 16, 16, 16, 1)
    p.decomp_2d_fft_set_ngrid(4)  # growing keeps the engines that exist (move_alloc, :388-398)
    assert p.decomp_2d_fft_get_engine(1) is eng and p.decomp_2d_fft_get_ngrid() == 4
    a = d2d.alloc_x(torch.complex64)
    b = d2d.alloc_z(torch.complex64)
    with pytest.raises(p.Decomp2dError, match="engine precision"):
        eng.fft_3d(a, b, p.DECOMP_2D_FFT_FORWARD)  # a float64 engine must not be handed complex64 arrays
    c = d2d.alloc_x(torch.complex128)
    with pytest.raises(p.Decomp2dError, match="overlap"):
        eng.fft_3d(c, c, p.DECOMP_2D_FFT_FORWARD)
    h = d2d.alloc_x(torch.float64, opt_levels=(1, 2, 0))
    assert tuple(h.shape) == (18, 20, 16) and h.lbound == (0, -1, 1)
    p.decomp_2d_finalize()


@pytest.mark.parametrize("shape", [(64, 32, 16), (256, 8, 512)])
def test_c2c_inplace_overwrites_input_and_matches(shape):
    """opt_inplace (src/fft_cufft.f90:696-706): the transform may use the input array as work space; the result is the same"""
    import torch
    p = pkg()
    d2d = p.decomp_2d_init(*shape, 1, 1)
    e0 = p.Decomp2dFFTEngine(d2d, p.PHYSICAL_IN_X)
    e1 = p.Decomp2dFFTEngine(d2d, p.PHYSICAL_IN_X, opt_inplace=True)
    assert e1.inplace
    a = d2d.alloc_x(torch.complex128)
    a.copy_(torch.complex(torch.rand(shape, dtype=torch.float64, device=a.device), torch.rand(shape, dtype=torch.float64, device=a.device)))
    a2 = a.clone().permute(2, 1, 0).contiguous().permute(2, 1, 0)
    o0, o1 = d2d.alloc_z(torch.complex128), d2d.alloc_z(torch.complex128)
    e0.fft_3d(a, o0, p.DECOMP_2D_FFT_FORWARD)
    e1.fft_3d(a2, o1, p.DECOMP_2D_FFT_FORWARD)
    assert torch.equal(o0, o1)
    back = d2d.alloc_x(torch.complex128)
    e1.fft_3d(o1, back, p.DECOMP_2D_FFT_BACKWARD)
    assert float((back / np.prod(shape) - a).abs().max().item()) < 1e-13
    e0.fin()
    e1.fin()
    d2d.finalize()
