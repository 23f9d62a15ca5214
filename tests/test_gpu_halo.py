"""Halo-cell exchange on the device (d2d_halo_update through the mirror's update_halo) against the numpy restatement of
update_halo / halo_exchange (src/halo.f90:101-198, 311-399; src/halo_exchange_{x,y,z}_body.f90): data movement, so every cell
must be bit-identical -- ghost layers from the face neighbours, the corners that the second exchange carries, periodic wrap
(one, two or several ranks along an axis) and untouched layers beyond non-periodic boundaries.  Multi-rank grids run as one
thread per rank on one GPU (d2d_group transport); tools/mgpu_check.py covers the peer-memory transport."""
import numpy as np
import pytest

import oracle as orc
from util import pkg, run_ranks

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("periodic", [(False, False, False), (True, True, True), (True, False, True)])
@pytest.mark.parametrize("level", [1, 2])
@pytest.mark.parametrize("pencil", [0, 1, 2])
@pytest.mark.parametrize("grid", [(1, 1), (1, 2), (2, 1), (2, 2), (2, 4), (3, 2)])
@pytest.mark.parametrize("shape,dtype", [((17, 13, 11), np.float64), ((16, 16, 16), np.complex64)])
def test_update_halo_bit_exact(shape, dtype, grid, pencil, level, periodic):
    import torch
    p = pkg()
    m = np.arange(1, np.prod(shape) + 1, dtype=np.float64).reshape(shape, order="F")
    g = (m + 1j * (m - 0.5)).astype(dtype) if np.dtype(dtype).kind == "c" else m.astype(dtype)
    want = orc.update_halo_world(g, grid, pencil, level, periodic)
    ins = orc.scatter(g, grid, pencil)
    tdt = {np.float64: torch.float64, np.complex64: torch.complex64}[dtype]
    nranks = grid[0] * grid[1]

    def body(rank, group):
        d2d = p.Decomp2d(*shape, grid[0], grid[1], rank=rank, nranks=nranks, group=group, device=0, periodic_bc=periodic)
        a = (d2d.alloc_x, d2d.alloc_y, d2d.alloc_z)[pencil](tdt)
        a.copy_(torch.from_numpy(ins[rank]))
        out = d2d.update_halo(a, level, opt_pencil=pencil + 1, opt_global=True)
        st = (d2d.decomp_main.xst, d2d.decomp_main.yst, d2d.decomp_main.zst)[pencil]
        assert out.lbound == tuple(st[i] - (0 if i == pencil else level) for i in range(3))
        got = out.cpu().numpy()
        again = d2d.update_halo(a, level, opt_pencil=pencil + 1).cpu().numpy()  # work buffers reused
        d2d.finalize()
        return np.array_equal(got, want[rank]), np.array_equal(again, want[rank])

    res = run_ranks(nranks, body) if nranks > 1 else [body(0, None)]
    assert all(a and b for a, b in res), res


def test_update_halo_pencil_deduced_from_shape():
    """without opt_pencil the pencil is deduced from the array shape, X first (src/halo.f90:201-245, the deprecated interface)"""
    import torch
    p = pkg()
    d2d = p.Decomp2d(8, 6, 4, 1, 1, periodic_bc=(True, True, True))
    a = d2d.alloc_y(torch.float64)
    a.copy_(torch.arange(1, 8 * 6 * 4 + 1, dtype=torch.float64, device=a.device).reshape(4, 6, 8).permute(2, 1, 0))
    out = d2d.update_halo(a, 1)  # on a 1 x 1 grid every pencil has the global shape: treated as an X-pencil
    assert tuple(out.shape) == (8, 8, 6)
    g = a.cpu().numpy()
    want = np.pad(g, ((0, 0), (1, 1), (1, 1)), mode="wrap")
    assert np.array_equal(out.cpu().numpy(), want)
    with pytest.raises(p.Decomp2dError, match="Invalid data passed to update_halo"):
        d2d.update_halo(a, 1, opt_pencil=4)
    d2d.finalize()
