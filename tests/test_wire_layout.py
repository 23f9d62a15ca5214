"""CPU emulation of the private wire layouts used inside decomp_2d_fft_3d (no GPU): a producer
stage writes its pencil through its link map, the all-to-all-v moves the blocks with the reference's
counts / displacements, the consumer stage gathers through its link map -- and must see exactly the
destination pencil of the reference transpose (oracle), for every link direction, ragged grids included."""
import ctypes as C

import numpy as np
import pytest

import oracle as orc
from util import pkg

AXES = {0: (0, 1, 2), 1: (1, 2, 0), 2: (2, 0, 1)}  # stage pencil -> (axis of e, axis of a, axis of b)


def link_map(p, dec, pencil, other, consumer):
    lib = p.lib()
    np_, na, nb = C.c_int(), C.c_int(), C.c_int()
    e0 = (C.c_int * 9)()
    off, se, sa, sb = [(C.c_int64 * 8)() for _ in range(4)]
    in_self = (C.c_int * 8)()
    rc = lib.d2d_debug_link_map(dec._h, pencil, other, consumer, C.byref(np_), e0, off, in_self, se, sa, sb, C.byref(na), C.byref(nb))
    assert rc == 0, lib.d2d_last_error()
    n = np_.value
    return dict(np=n, e0=list(e0)[: n + 1], off=list(off)[:n], in_self=list(in_self)[:n], se=list(se)[:n], sa=list(sa)[:n],
                sb=list(sb)[:n], na=na.value, nb=nb.value)


def offsets(m, e, a, b):
    """element offsets and buffer selector for index arrays e, a, b"""
    piece = np.searchsorted(np.array(m["e0"][1:]), e, side="right")
    piece = np.minimum(piece, m["np"] - 1)
    off = np.array(m["off"])[piece] + (e - np.array(m["e0"])[piece]) * np.array(m["se"])[piece] + a * np.array(m["sa"])[piece] \
        + b * np.array(m["sb"])[piece]
    return off, np.array(m["in_self"])[piece]


@pytest.mark.parametrize("grid", [(1, 1), (1, 2), (2, 1), (2, 2), (2, 4), (3, 2), (4, 2)])
@pytest.mark.parametrize("shape", [(17, 13, 11), (32, 16, 33), (9, 24, 16)])
@pytest.mark.parametrize("link", [(0, 1), (1, 0), (1, 2), (2, 1)])
def test_link_maps_compose_to_the_reference_transpose(shape, grid, link):
    p = pkg()
    P, Q = link
    nranks = grid[0] * grid[1]
    nx, ny, nz = shape
    g = np.arange(nx * ny * nz, dtype=np.float64).reshape(shape, order="F")
    src = orc.scatter(g, grid, P)
    want = orc.scatter(g, grid, Q)
    decs = [p.DecompInfo.for_rank(nx, ny, nz, grid[0], grid[1], r) for r in range(nranks)]
    col = (P == 0 or Q == 0)

    def tables(d, pen, oth):
        if pen == 0:
            return d.x1cnts, d.x1disp
        if pen == 2:
            return d.z2cnts, d.z2disp
        return (d.y1cnts, d.y1disp) if oth == 0 else (d.y2cnts, d.y2disp)

    sendbufs, recvbufs = [], []
    for r in range(nranks):
        m = link_map(p, decs[r], P, Q, 0)
        sz = (decs[r].xsz, decs[r].ysz, decs[r].zsz)[P]
        ax = AXES[P]
        assert (m["na"], m["nb"]) == (sz[ax[1]], sz[ax[2]])
        idx = np.indices(sz)
        e, a, b = idx[ax[0]].ravel(), idx[ax[1]].ravel(), idx[ax[2]].ravel()
        off, sel = offsets(m, e, a, b)
        assert np.all(sel == (np.searchsorted(np.array(m["e0"][1:]), e, side="right") == (r // grid[1] if col else r % grid[1])))
        buf = np.full(int(np.prod(sz)), np.nan)
        assert len(np.unique(off)) == len(off) and off.min() == 0 and off.max() == len(buf) - 1, "producer map must be a bijection"
        buf[off] = src[r].ravel(order="C")[np.ravel_multi_index((idx[0].ravel(), idx[1].ravel(), idx[2].ravel()), sz)]
        sendbufs.append(buf)
        recvbufs.append(np.full(int(np.prod((decs[r].xsz, decs[r].ysz, decs[r].zsz)[Q])), np.nan))
    # all-to-all-v with the reference's counts / displacements (self block excluded: it never moves)
    for r in range(nranks):
        c1, c2 = r // grid[1], r % grid[1]
        me = c1 if col else c2
        rc, rd = tables(decs[r], Q, P)
        for mth in range(grid[0] if col else grid[1]):
            if mth == me:
                continue
            peer = (mth * grid[1] + c2) if col else (c1 * grid[1] + mth)
            sc, sd = tables(decs[peer], P, Q)
            assert sc[me] == rc[mth]
            recvbufs[r][rd[mth]: rd[mth] + rc[mth]] = sendbufs[peer][sd[me]: sd[me] + sc[me]]
    for r in range(nranks):
        m = link_map(p, decs[r], Q, P, 1)
        sz = (decs[r].xsz, decs[r].ysz, decs[r].zsz)[Q]
        ax = AXES[Q]
        idx = np.indices(sz)
        e, a, b = idx[ax[0]].ravel(), idx[ax[1]].ravel(), idx[ax[2]].ravel()
        off, sel = offsets(m, e, a, b)
        vals = np.where(sel == 1, sendbufs[r][np.minimum(off, len(sendbufs[r]) - 1)], recvbufs[r][np.minimum(off, len(recvbufs[r]) - 1)])
        got = np.zeros(sz)
        got[idx[0].ravel(), idx[1].ravel(), idx[2].ravel()] = vals
        assert np.array_equal(got, want[r]), (link, r)
    for d in decs:
        d.finalize()
