"""The host-array entry points d2d_fft_3d_{r2c,c2r,c2c}_host (what the reference's examples do with `!$acc data copyin(in) copy(out)`
around decomp_2d_fft_3d, examples/fft_physical_z/fft_r2c_z.f90; bench.py's e2e path): upload, transform, download as three pipelines
on three streams, z-slab gating of x stages, per-piece tracking of host-memory dependencies between calls.  Checked against the
oracle in blocking mode and in stream-ordered mode (calls return once enqueued; d2d_ctx_sync completes them) with the SAME host
buffers reused call after call, pinned and pageable host memory, both formats and precisions, one and several ranks."""
import numpy as np
import pytest

import oracle as orc
from util import pkg, run_ranks

pytestmark = pytest.mark.gpu


def _host(torch, arr, pinned):
    t = torch.from_numpy(np.ascontiguousarray(arr.reshape(-1, order="F").view(np.float64 if arr.dtype in (np.float64, np.complex128) else np.float32)))
    return t.pin_memory() if pinned else t.clone()


@pytest.mark.parametrize("pinned", [True, False])
@pytest.mark.parametrize("blocking", [True, False])
@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("fmt", [orc.PHYSICAL_IN_X, orc.PHYSICAL_IN_Z])
@pytest.mark.parametrize("shape", [(64, 32, 128), (256, 24, 512), (34, 26, 22)])
def test_host_arrays_chained_calls(shape, fmt, prec, blocking, pinned):
    import torch
    p = pkg()
    rdt, cdt = (np.float64, np.complex128) if prec == "f64" else (np.float32, np.complex64)
    tol = 1e-12 if prec == "f64" else 1e-5
    rng = np.random.default_rng(3)
    fields = [np.asfortranarray(rng.uniform(-1, 1, shape)).astype(rdt) for _ in range(3)]
    refs = [orc.fft_3d_r2c_world(shape, (1, 1), fmt, [g.astype(np.float64)])[0] for g in fields]
    d2d = p.decomp_2d_init(*shape, 1, 1)
    d2d.set_blocking(blocking)
    eng = p.decomp_2d_fft_init(fmt, dtype=torch.float64 if prec == "f64" else torch.float32)
    h_in = [_host(torch, g, pinned) for g in fields]
    h_spec = _host(torch, np.zeros(refs[0].shape, dtype=cdt, order="F"), pinned)   # ONE spectrum buffer, reused by every pair
    h_back = [_host(torch, np.zeros(shape, dtype=rdt, order="F"), pinned) for _ in fields]
    specs = []
    for i in range(3):
        eng.fft_3d_r2c_host(h_in[i].data_ptr(), h_spec.data_ptr())
        if i == 1:  # look at one spectrum on the way (forces completion in stream-ordered mode)
            d2d.sync()
            specs.append(h_spec.numpy().copy())
        eng.fft_3d_c2r_host(h_spec.data_ptr(), h_back[i].data_ptr())
    d2d.sync()
    got = specs[0].view(cdt).reshape(refs[1].shape, order="F")
    assert np.max(np.abs(got - refs[1])) / np.max(np.abs(refs[1])) < tol
    for i in range(3):
        back = h_back[i].numpy().reshape(shape, order="F") / np.prod(shape)
        assert np.max(np.abs(back - fields[i])) < (1e-13 if prec == "f64" else 2e-5), i
    # the last spectrum is still in the shared buffer
    last = h_spec.numpy().view(cdt).reshape(refs[2].shape, order="F")
    assert np.max(np.abs(last - refs[2])) / np.max(np.abs(refs[2])) < tol
    p.decomp_2d_finalize()


@pytest.mark.parametrize("blocking", [True, False])
@pytest.mark.parametrize("fmt", [orc.PHYSICAL_IN_X, orc.PHYSICAL_IN_Z])
def test_host_arrays_c2c_and_in_place_reuse(fmt, blocking):
    """c2c through host arrays, forward then backward INTO THE INPUT ARRAY of the forward call (write after read on the host)"""
    import torch
    p = pkg()
    shape = (32, 64, 16)
    rng = np.random.default_rng(5)
    g = np.asfortranarray(rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape))
    pin = 0 if fmt == orc.PHYSICAL_IN_X else 2
    ref = orc.fft_3d_c2c_world(shape, (1, 1), fmt, orc.FORWARD, orc.scatter(g, (1, 1), pin))[0]
    d2d = p.decomp_2d_init(*shape, 1, 1)
    d2d.set_blocking(blocking)
    eng = p.decomp_2d_fft_init(fmt)
    h_a = _host(torch, g, True)
    h_b = _host(torch, np.zeros(shape, dtype=np.complex128, order="F"), True)
    eng.fft_3d_c2c_host(h_a.data_ptr(), h_b.data_ptr(), p.DECOMP_2D_FFT_FORWARD)
    eng.fft_3d_c2c_host(h_b.data_ptr(), h_a.data_ptr(), p.DECOMP_2D_FFT_BACKWARD)
    d2d.sync()
    spec = h_b.numpy().view(np.complex128).reshape(ref.shape, order="F")
    assert np.max(np.abs(spec - ref)) / np.max(np.abs(ref)) < 1e-12
    back = h_a.numpy().view(np.complex128).reshape(shape, order="F") / np.prod(shape)
    assert np.max(np.abs(back - g)) < 1e-13
    p.decomp_2d_finalize()


@pytest.mark.parametrize("grid", [(1, 2), (2, 2)])
@pytest.mark.parametrize("fmt", [orc.PHYSICAL_IN_X, orc.PHYSICAL_IN_Z])
def test_host_arrays_multi_rank(grid, fmt):
    import torch
    p = pkg()
    shape = (64, 32, 48)
    rng = np.random.default_rng(9)
    g = np.asfortranarray(rng.uniform(-1, 1, shape))
    pin = 0 if fmt == orc.PHYSICAL_IN_X else 2
    ins = orc.scatter(g, grid, pin)
    ref = orc.fft_3d_r2c_world(shape, grid, fmt, ins)
    nranks = grid[0] * grid[1]

    def body(rank, group):
        d2d = p.Decomp2d(*shape, grid[0], grid[1], rank=rank, nranks=nranks, group=group, device=0)
        d2d.set_blocking(False)
        eng = p.Decomp2dFFTEngine(d2d, fmt)
        h_in = _host(torch, ins[rank], True)
        h_spec = _host(torch, np.zeros(ref[rank].shape, dtype=np.complex128, order="F"), True)
        h_back = _host(torch, np.zeros(ins[rank].shape, dtype=np.float64, order="F"), True)
        for _ in range(2):
            eng.fft_3d_r2c_host(h_in.data_ptr(), h_spec.data_ptr())
            eng.fft_3d_c2r_host(h_spec.data_ptr(), h_back.data_ptr())
        d2d.sync()
        out = (h_spec.numpy().view(np.complex128).reshape(ref[rank].shape, order="F").copy(),
               h_back.numpy().reshape(ins[rank].shape, order="F").copy() / np.prod(shape))
        eng.fin()
        d2d.finalize()
        return out

    res = run_ranks(nranks, body)
    smax = max(np.max(np.abs(s)) for s in ref)
    for r in range(nranks):
        if ref[r].size:
            assert np.max(np.abs(res[r][0] - ref[r])) / smax < 1e-12, r
        assert np.max(np.abs(res[r][1] - ins[r])) < 1e-13, r
