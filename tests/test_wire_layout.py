"""CPU emulation of the private wire layouts used inside decomp_2d_fft_3d (no GPU): a producer
stage writes its pencil through its link map, the all-to-all-v moves the (padded) blocks with the
library's own counts / displacements, the consumer stage gathers through its link map -- and must
see exactly the destination pencil of the reference transpose (oracle), for every link direction,
ragged grids and both paddings (fp64: 8 elements = 128 B, fp32: 16 elements) included."""
import ctypes as C

import numpy as np
import pytest

import oracle as orc
from util import pkg

AXES = {0: (0, 1, 2), 1: (1, 2, 0), 2: (2, 0, 1)}  # stage pencil -> (axis of e, axis of a, axis of b)


def link_map(p, dec, pencil, other, consumer, padq):
    lib = p.lib()
    np_, na, nb = C.c_int(), C.c_int(), C.c_int()
    e0 = (C.c_int * 9)()
    off, se, sa, sb, cnt, disp = [(C.c_int64 * 8)() for _ in range(6)]
    in_self = (C.c_int * 8)()
    rc = lib.d2d_debug_link_map(dec._h, pencil, other, consumer, padq, C.byref(np_), e0, off, in_self, se, sa, sb, C.byref(na),
                                C.byref(nb), cnt, disp)
    assert rc == 0, lib.d2d_last_error()
    n = np_.value
    return dict(np=n, e0=list(e0)[: n + 1], off=list(off)[:n], in_self=list(in_self)[:n], se=list(se)[:n], sa=list(sa)[:n],
                sb=list(sb)[:n], na=na.value, nb=nb.value, cnt=list(cnt)[:n], disp=list(disp)[:n])


def offsets(m, e, a, b):
    """element offsets and buffer selector for index arrays e, a, b"""
    piece = np.searchsorted(np.array(m["e0"][1:]), e, side="right")
    piece = np.minimum(piece, m["np"] - 1)
    off = np.array(m["off"])[piece] + (e - np.array(m["e0"])[piece]) * np.array(m["se"])[piece] + a * np.array(m["sa"])[piece] \
        + b * np.array(m["sb"])[piece]
    return off, np.array(m["in_self"])[piece]


@pytest.mark.parametrize("padq", [8, 16, 1])
@pytest.mark.parametrize("grid", [(1, 1), (1, 2), (2, 1), (2, 2), (2, 4), (3, 2), (4, 2), (8, 1), (1, 8)])
@pytest.mark.parametrize("shape", [(17, 13, 11), (32, 16, 33), (9, 24, 16)])
@pytest.mark.parametrize("link", [(0, 1), (1, 0), (1, 2), (2, 1)])
def test_link_maps_compose_to_the_reference_transpose(shape, grid, link, padq):
    p = pkg()
    P, Q = link
    nranks = grid[0] * grid[1]
    nx, ny, nz = shape
    g = np.arange(nx * ny * nz, dtype=np.float64).reshape(shape, order="F")
    src = orc.scatter(g, grid, P)
    want = orc.scatter(g, grid, Q)
    decs = [p.DecompInfo.for_rank(nx, ny, nz, grid[0], grid[1], r) for r in range(nranks)]
    col = (P == 0 or Q == 0)
    prod = [link_map(p, decs[r], P, Q, 0, padq) for r in range(nranks)]
    cons = [link_map(p, decs[r], Q, P, 1, padq) for r in range(nranks)]
    sendbufs, recvbufs = [], []
    for r in range(nranks):
        m = prod[r]
        sz = (decs[r].xsz, decs[r].ysz, decs[r].zsz)[P]
        ax = AXES[P]
        assert (m["na"], m["nb"]) == (sz[ax[1]], sz[ax[2]])
        idx = np.indices(sz)
        e, a, b = idx[ax[0]].ravel(), idx[ax[1]].ravel(), idx[ax[2]].ravel()
        off, sel = offsets(m, e, a, b)
        total = m["disp"][-1] + m["cnt"][-1]
        assert total >= int(np.prod(sz)) and (padq > 1 or total == int(np.prod(sz)))
        assert len(np.unique(off)) == len(off) and off.min() >= 0 and off.max() < total, "producer map must be injective"
        # rows of every block start on a padq boundary
        rows = off[e == np.array(m["e0"])[np.minimum(np.searchsorted(np.array(m["e0"][1:]), e, side="right"), m["np"] - 1)]]
        if padq > 1 and min(m["se"]) == 1:
            assert np.all(rows % padq == 0)
        buf = np.full(total, np.nan)
        buf[off] = src[r][idx[0].ravel(), idx[1].ravel(), idx[2].ravel()]
        sendbufs.append(buf)
        recvbufs.append(np.full(cons[r]["disp"][-1] + cons[r]["cnt"][-1], np.nan))
    # all-to-all-v with the library's counts / displacements (self block excluded: it never moves)
    for r in range(nranks):
        c1, c2 = r // grid[1], r % grid[1]
        me = c1 if col else c2
        for mth in range(grid[0] if col else grid[1]):
            if mth == me:
                continue
            peer = (mth * grid[1] + c2) if col else (c1 * grid[1] + mth)
            sc, sd = prod[peer]["cnt"][me], prod[peer]["disp"][me]
            rc, rd = cons[r]["cnt"][mth], cons[r]["disp"][mth]
            assert sc == rc, "send / recv block sizes must agree"
            recvbufs[r][rd: rd + rc] = sendbufs[peer][sd: sd + sc]
    for r in range(nranks):
        m = cons[r]
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


def link_chunk(p, dec, pencil, other, padq, f0, f1):
    lib = p.lib()
    np_, axis_is_a, nf = C.c_int(), C.c_int(), C.c_int()
    off, cnt = (C.c_int64 * 8)(), (C.c_int64 * 8)()
    rc = lib.d2d_debug_link_chunk(dec._h, pencil, other, padq, f0, f1, C.byref(np_), C.byref(axis_is_a), C.byref(nf), off, cnt)
    assert rc == 0, lib.d2d_last_error()
    return dict(np=np_.value, axis_is_a=axis_is_a.value, nf=nf.value, off=list(off)[: np_.value], cnt=list(cnt)[: np_.value])


@pytest.mark.parametrize("nchunks", [1, 3])
@pytest.mark.parametrize("grid", [(1, 2), (2, 1), (2, 2), (2, 4), (3, 2)])
@pytest.mark.parametrize("shape", [(17, 13, 11), (32, 16, 33)])
@pytest.mark.parametrize("link", [(0, 1), (1, 0), (1, 2), (2, 1)])
def test_chunked_exchange_pipeline_moves_exactly_what_each_chunk_needs(shape, grid, link, nchunks):
    """Chunk-wise overlap of an exchange with its neighbouring stages, emulated on the CPU: the link is cut
    along its free axis; for every chunk the producer writes only the sub-ranges d2d_debug_link_chunk announces, the exchange
    moves only those sub-ranges, and the consumer restricted to the same chunk already sees its part of the reference pencil --
    before any later chunk has been produced."""
    p = pkg()
    P, Q = link
    padq = 8
    nranks = grid[0] * grid[1]
    nx, ny, nz = shape
    g = np.arange(nx * ny * nz, dtype=np.float64).reshape(shape, order="F")
    src = orc.scatter(g, grid, P)
    want = orc.scatter(g, grid, Q)
    decs = [p.DecompInfo.for_rank(nx, ny, nz, grid[0], grid[1], r) for r in range(nranks)]
    col = (P == 0 or Q == 0)
    prod = [link_map(p, decs[r], P, Q, 0, padq) for r in range(nranks)]
    cons = [link_map(p, decs[r], Q, P, 1, padq) for r in range(nranks)]
    sendbufs = [np.full(prod[r]["disp"][-1] + prod[r]["cnt"][-1], np.nan) for r in range(nranks)]
    recvbufs = [np.full(cons[r]["disp"][-1] + cons[r]["cnt"][-1], np.nan) for r in range(nranks)]
    got = [np.full((decs[r].xsz, decs[r].ysz, decs[r].zsz)[Q], np.nan) for r in range(nranks)]

    def free_index(m_axis_is_a, a, b):
        return a if m_axis_is_a else b

    for c in range(nchunks):
        pch, cch = [], []
        for r in range(nranks):
            nf = link_chunk(p, decs[r], P, Q, padq, 0, 0)["nf"]
            f0, f1 = nf * c // nchunks, nf * (c + 1) // nchunks
            pch.append(link_chunk(p, decs[r], P, Q, padq, f0, f1))
            cc = link_chunk(p, decs[r], Q, P, padq, f0, f1)
            assert cc["nf"] == nf, "the free axis has the same extent on both sides of the link"
            cch.append(cc)
            pch[-1]["f"] = cch[-1]["f"] = (f0, f1)
        # producer stage, restricted to the chunk
        for r in range(nranks):
            m, ch = prod[r], pch[r]
            sz = (decs[r].xsz, decs[r].ysz, decs[r].zsz)[P]
            ax = AXES[P]
            idx = np.indices(sz)
            e, a, b = idx[ax[0]].ravel(), idx[ax[1]].ravel(), idx[ax[2]].ravel()
            f = free_index(ch["axis_is_a"], a, b)
            sel = (f >= ch["f"][0]) & (f < ch["f"][1])
            off, _ = offsets(m, e[sel], a[sel], b[sel])
            inside = np.zeros(len(off), dtype=bool)
            for k in range(ch["np"]):
                inside |= (off >= ch["off"][k]) & (off < ch["off"][k] + ch["cnt"][k])
            assert inside.all(), "a producer chunk writes only inside the announced sub-ranges"
            sendbufs[r][off] = src[r][idx[0].ravel()[sel], idx[1].ravel()[sel], idx[2].ravel()[sel]]
        # exchange of the chunk's sub-ranges only
        for r in range(nranks):
            c1, c2 = r // grid[1], r % grid[1]
            me = c1 if col else c2
            for mth in range(grid[0] if col else grid[1]):
                if mth == me:
                    continue
                peer = (mth * grid[1] + c2) if col else (c1 * grid[1] + mth)
                so, sc = pch[peer]["off"][me], pch[peer]["cnt"][me]
                ro, rc = cch[r]["off"][mth], cch[r]["cnt"][mth]
                assert sc == rc, "chunk sub-ranges of the two sides must have the same size"
                recvbufs[r][ro: ro + rc] = sendbufs[peer][so: so + sc]
        # consumer stage, restricted to the same chunk: everything it needs has arrived
        for r in range(nranks):
            m, ch = cons[r], cch[r]
            sz = (decs[r].xsz, decs[r].ysz, decs[r].zsz)[Q]
            ax = AXES[Q]
            idx = np.indices(sz)
            e, a, b = idx[ax[0]].ravel(), idx[ax[1]].ravel(), idx[ax[2]].ravel()
            f = free_index(ch["axis_is_a"], a, b)
            sel = (f >= ch["f"][0]) & (f < ch["f"][1])
            off, which = offsets(m, e[sel], a[sel], b[sel])
            vals = np.where(which == 1, sendbufs[r][np.minimum(off, len(sendbufs[r]) - 1)], recvbufs[r][np.minimum(off, len(recvbufs[r]) - 1)])
            i0, i1, i2 = idx[0].ravel()[sel], idx[1].ravel()[sel], idx[2].ravel()[sel]
            assert not np.isnan(vals).any(), "a consumer chunk must not depend on data of another chunk"
            assert np.array_equal(vals, want[r][i0, i1, i2]), (link, r, c)
            got[r][i0, i1, i2] = vals
    for r in range(nranks):
        assert np.array_equal(got[r], want[r])
    for d in decs:
        d.finalize()
