"""BASELINE.json `configs` as parity cases on ONE GPU (multi-rank grids run thread-per-rank through d2d_group:
same piece maps, pack/unpack and wire layouts as the NCCL path), plus size-independent properties at the full
sizes the oracle cannot reach.

  configs[0]  fft_physical_x 64^3 fp64 c2c forward+backward, p_row x p_col = 1 x 2  (examples/fft_physical_x/fft_c2c_x.f90)
  configs[1]  test2d transpose round trip 256^3 fp64 on 2 x 2, bit-exact               (examples/test2d/test2d.f90:92-199)
  configs[2]  fft_physical_x 512^3 fp64 r2c/c2r, 2 x 4                                 (examples/fft_physical_x/fft_r2c_x.f90)
  configs[3]  fft_physical_z 1024^3 fp64 r2c/c2r (headline), here 1 x 1: round trip + Parseval
  configs[4]  2048^3 fp32 r2c/c2r: round trip + Parseval on 1 x 1 when the GPU has the memory for it
"""
import numpy as np
import pytest

import oracle as orc
from util import pkg, run_ranks

pytestmark = pytest.mark.gpu


def test_config0_c2c_64_1x2_reference_field():
    import torch
    p = pkg()
    shape, grid = (64, 64, 64), (1, 2)
    nx, ny, nz = shape
    i, j, k = np.meshgrid(np.arange(1, nx + 1), np.arange(1, ny + 1), np.arange(1, nz + 1), indexing="ij")
    # the example's field: (1 + i) * (i/nx)(j/ny)(k/nz) (fft_c2c_x.f90:80-92)
    g = np.asfortranarray((i / nx) * (j / ny) * (k / nz) * (1 + 1j))
    ins = orc.scatter(g, grid, 0)
    ref = orc.fft_3d_c2c_world(shape, grid, orc.PHYSICAL_IN_X, orc.FORWARD, ins)

    def body(rank, group):
        d2d = p.Decomp2d(*shape, *grid, rank=rank, nranks=2, group=group, device=0)
        eng = p.Decomp2dFFTEngine(d2d, p.PHYSICAL_IN_X)
        a_in, a_out = d2d.alloc_x(torch.complex128), d2d.alloc_z(torch.complex128)
        a_in.copy_(torch.from_numpy(ins[rank]))
        eng.fft_3d(a_in, a_out, p.DECOMP_2D_FFT_FORWARD)
        spec = a_out.cpu().numpy()
        back = d2d.alloc_x(torch.complex128)
        eng.fft_3d(a_out, back, p.DECOMP_2D_FFT_BACKWARD)
        out = (spec, back.cpu().numpy())
        eng.fin()
        d2d.finalize()
        return out

    res = run_ranks(2, body)
    smax = max(np.max(np.abs(s)) for s in ref)
    for r in range(2):
        assert np.max(np.abs(res[r][0] - ref[r])) / smax < 1e-12
    rt = orc.gather([x[1] for x in res], shape, grid, 0) / np.prod(shape)
    # the example's own criterion: summed error per point <= eps * 50 (fft_c2c_x.f90:140-154)
    assert np.sum(np.abs(rt - g)) / np.prod(shape) < np.finfo(np.float64).eps * 50


@pytest.mark.parametrize("cplx", [False, True])
def test_config1_test2d_256_2x2_bit_exact(cplx):
    import torch
    p = pkg()
    shape, grid = (256, 256, 256), (2, 2)
    nx, ny, nz = shape
    m = (np.arange(1, nx * ny * nz + 1, dtype=np.float64)).reshape(shape, order="F")  # test2d's index field
    g = np.asfortranarray(m + 1j * (m - 1)) if cplx else m
    want = [orc.scatter(g, grid, pen) for pen in range(3)]
    tdt = torch.complex128 if cplx else torch.float64

    def body(rank, group):
        d2d = p.Decomp2d(*shape, *grid, rank=rank, nranks=4, group=group, device=0)
        u1, u2, u3 = d2d.alloc_x(tdt), d2d.alloc_y(tdt), d2d.alloc_z(tdt)
        u1.copy_(torch.from_numpy(want[0][rank]))
        d2d.transpose_x_to_y(u1, u2)
        ok = [np.array_equal(u2.cpu().numpy(), want[1][rank])]
        d2d.transpose_y_to_z(u2, u3)
        ok.append(np.array_equal(u3.cpu().numpy(), want[2][rank]))
        u2.zero_()
        d2d.transpose_z_to_y(u3, u2)
        ok.append(np.array_equal(u2.cpu().numpy(), want[1][rank]))
        u1.zero_()
        d2d.transpose_y_to_x(u2, u1)
        ok.append(np.array_equal(u1.cpu().numpy(), want[0][rank]))
        d2d.finalize()
        return ok

    for r, ok in enumerate(run_ranks(4, body)):
        assert ok == [True] * 4, (r, ok)


def test_config2_r2c_512_2x4_physical_in_x():
    import torch
    p = pkg()
    shape, grid = (512, 512, 512), (2, 4)
    rng = np.random.default_rng(20240601 + 2)
    g = np.asfortranarray(rng.uniform(-1, 1, shape))
    ins = orc.scatter(g, grid, 0)
    ref = orc.fft_3d_r2c_world(shape, grid, orc.PHYSICAL_IN_X, ins)  # sp is ragged: x1dist = (128, 129)

    def body(rank, group):
        d2d = p.Decomp2d(*shape, *grid, rank=rank, nranks=8, group=group, device=0)
        eng = p.Decomp2dFFTEngine(d2d, p.PHYSICAL_IN_X)
        in_r, out_c = d2d.alloc_x(torch.float64, eng.ph), d2d.alloc_z(torch.complex128, eng.sp)
        in_r.copy_(torch.from_numpy(ins[rank]))
        eng.fft_3d(in_r, out_c)
        spec = out_c.cpu().numpy()
        back = d2d.alloc_x(torch.float64, eng.ph)
        eng.fft_3d(out_c, back)
        out = (spec, back.cpu().numpy())
        eng.fin()
        d2d.finalize()
        return out

    res = run_ranks(8, body, timeout=600)
    smax = max(np.max(np.abs(s)) for s in ref)
    for r in range(8):
        assert np.max(np.abs(res[r][0] - ref[r])) / smax < 1e-12, r
    rt = orc.gather([x[1] for x in res], shape, grid, 0) / np.prod(shape)
    assert np.max(np.abs(rt - g)) < 1e-13


def _parseval_half(spec, axis, n_full, planes=16):
    """sum |X|^2 over the full spectrum from the Hermitian half along `axis` (bins 0..n/2); z-slabs keep temporaries small"""
    import torch
    w = torch.full((spec.shape[axis],), 2.0, dtype=torch.float64, device=spec.device)
    w[0] = 1.0
    if n_full % 2 == 0:
        w[-1] = 1.0
    total = 0.0
    for z0 in range(0, spec.shape[2], planes):
        c = spec[:, :, z0:z0 + planes]
        e = c.real.double() ** 2 + c.imag.double() ** 2
        if axis == 2:
            total += float((e.sum(dim=(0, 1)) * w[z0:z0 + planes]).sum().item())
        else:
            total += float((e.sum(dim=(1, 2)) * w).sum().item())
    return total


def _max_abs_diff(a, b, planes=16):
    return max(float((a[:, :, z0:z0 + planes] - b[:, :, z0:z0 + planes]).abs().max().item()) for z0 in range(0, a.shape[2], planes))


@pytest.mark.parametrize("case", ["1024_f64_z", "2048_f32_x"])
def test_full_size_properties(case):
    """configs[3] / configs[4] at full size on a 1 x 1 grid: the forward spectrum of a separable seeded field against the
    oracle's 1-D transforms of its factor lines (all bins), then round trip and Parseval of a random field;
    c2r writes back into the input array and the input is regenerated from its seed for the comparison."""
    import torch
    p = pkg()
    n, prec, fmt = (1024, "f64", p.PHYSICAL_IN_Z) if case.startswith("1024") else (2048, "f32", p.PHYSICAL_IN_X)
    rdt, cdt = (torch.float64, torch.complex128) if prec == "f64" else (torch.float32, torch.complex64)
    rs = 8 if prec == "f64" else 4
    need = 4.2 * rs * n ** 3 + 6 * 2 ** 30  # input + spectrum + two work buffers (each about one pencil) + slab temporaries
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    if free < need:
        pytest.skip(f"needs {need / 2**30:.0f} GiB of device memory, {free / 2**30:.0f} free")
    d2d = p.decomp_2d_init(n, n, n, 1, 1)
    eng = p.decomp_2d_fft_init(fmt, dtype=rdt)
    alloc_in, alloc_out = (d2d.alloc_z, d2d.alloc_x) if fmt == p.PHYSICAL_IN_Z else (d2d.alloc_x, d2d.alloc_z)
    a = alloc_in(rdt, eng.ph)
    spec = alloc_out(cdt, eng.sp)
    # (1) forward spectrum against ORACLE lines: a separable random field f = fa(i) fb(j) fc(k) has the spectrum
    #     A(kx) B(ky) C(kz), where A, B, C are the oracle's (Glassman) 1-D transforms of the three seeded factor lines --
    #     every bin of the full-size spectrum is compared, at the cost of three oracle lines
    rng = np.random.default_rng(20240601)
    fac = [rng.uniform(-1, 1, n) for _ in range(3)]
    half_axis = 2 if fmt == p.PHYSICAL_IN_Z else 0
    lines = []
    for ax in range(3):
        col = np.asfortranarray(fac[ax].reshape(n, 1, 1))
        if ax == half_axis:
            lines.append(orc.r2c_1m(col, 0).reshape(-1))
        else:
            lines.append(orc.c2c_1m(col.astype(np.complex128), 0, orc.FORWARD).reshape(-1))
    dev = a.device
    tf = [torch.from_numpy(f).to(dev) for f in fac]
    plane = tf[0][:, None] * tf[1][None, :]
    for k in range(n):
        a[:, :, k] = (plane * tf[2][k]).to(rdt)
    eng.fft_3d(a, spec)
    tl = [torch.from_numpy(l).to(dev) for l in lines]
    refmax = float(tl[0].abs().max() * tl[1].abs().max() * tl[2].abs().max())
    rp = tl[0][:, None] * tl[1][None, :]
    worst = 0.0
    for k in range(spec.shape[2]):
        worst = max(worst, float((spec[:, :, k].to(torch.complex128) - rp * tl[2][k]).abs().max().item()))
    assert worst / refmax < (1e-12 if prec == "f64" else 1e-5), worst / refmax
    del plane, rp
    # (2) size-independent properties of a random field
    gen = torch.Generator(device=a.device)
    gen.manual_seed(20240601)
    a.uniform_(-1, 1, generator=gen)
    e_in = sum(float((a[:, :, z0:z0 + 16].double() ** 2).sum().item()) for z0 in range(0, a.shape[2], 16))  # slabs: small temporaries
    eng.fft_3d(a, spec)
    e_spec = _parseval_half(spec, 2 if fmt == p.PHYSICAL_IN_Z else 0, n)
    assert abs(e_spec / float(n) ** 3 - e_in) / e_in < (1e-12 if prec == "f64" else 1e-5)
    eng.fft_3d(spec, a)  # c2r back into the input array
    del spec
    torch.cuda.empty_cache()
    a.mul_(1.0 / float(n) ** 3)
    ref = torch.empty_like(a)
    gen.manual_seed(20240601)
    ref.uniform_(-1, 1, generator=gen)
    worst = _max_abs_diff(a, ref)
    assert worst < (1e-13 if prec == "f64" else 2e-5), worst
    del a, ref
    p.decomp_2d_finalize()
    torch.cuda.empty_cache()
