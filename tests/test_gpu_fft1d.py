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
    p, d2d, torch = _ctx()
    t = _falloc(torch, (3 * 5 * 7 * 11 * 13, 2, 2), torch.complex128)  # beyond the any-length kernel's shared-memory limit
    with pytest.raises(p.Decomp2dError, match="not supported"):
        d2d.c2c_1m(t, 0, -1)
    d2d.finalize()
