"""ctypes front-end of the CPU oracle (oracle/d2d_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product path (the CUDA library behind include/d2d_b200.h) never does.

All arrays are numpy, Fortran-ordered, one per simulated MPI rank ("world" = list indexed by rank,
rank r has coord (r // p_col, r % p_col) as in src/decomp_2d_init_fin.f90:95-123 of the reference).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_d2d.so")
MAXP = 64

X_TO_Y, Y_TO_Z, Z_TO_Y, Y_TO_X = 0, 1, 2, 3
PHYSICAL_IN_X, PHYSICAL_IN_Z = 1, 3
FORWARD, BACKWARD = -1, 1


def build(force=False):
    src = os.path.join(_HERE, "d2d_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class _Decomp(C.Structure):
    _fields_ = (
        [(n, C.c_int * 3) for n in ("xst", "xen", "xsz", "yst", "yen", "ysz", "zst", "zen", "zsz")]
        + [(n, C.c_int * MAXP) for n in ("x1dist", "y1dist", "y2dist", "z2dist")]
        + [(n, C.c_int64 * MAXP) for n in ("x1cnts", "y1cnts", "y2cnts", "z2cnts")]
        + [(n, C.c_int64 * MAXP) for n in ("x1disp", "y1disp", "y2disp", "z2disp")]
    )


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        assert _lib.orc_sizeof_decomp() == C.sizeof(_Decomp)
    return _lib


class Decomp:
    """decomp_info of the reference (src/info.f90:11-47), 0-based starts."""

    def __init__(self, nx, ny, nz, p_row, p_col, rank):
        d = _Decomp()
        lib().orc_decomp_init(nx, ny, nz, p_row, p_col, rank, C.byref(d))
        for n in ("xst", "xen", "xsz", "yst", "yen", "ysz", "zst", "zen", "zsz"):
            setattr(self, n, tuple(getattr(d, n)))
        for n in ("x1dist", "y1dist", "x1cnts", "y1cnts", "x1disp", "y1disp"):
            setattr(self, n, tuple(getattr(d, n))[:p_row])
        for n in ("y2dist", "z2dist", "y2cnts", "z2cnts", "y2disp", "z2disp"):
            setattr(self, n, tuple(getattr(d, n))[:p_col])
        self.shape = (nx, ny, nz)
        self.grid = (p_row, p_col)
        self.rank = rank

    def sz(self, pencil):
        return (self.xsz, self.ysz, self.zsz)[pencil]

    def st(self, pencil):
        return (self.xst, self.yst, self.zst)[pencil]


def set_threads(n):
    """OpenMP threads of the rank loops (one thread plays one MPI rank); returns the count now in effect."""
    lib().orc_set_threads(int(n))
    return int(lib().orc_get_max_threads())


def even_counts(nx, ny, nz, p_row, p_col, rank):
    """(x1count, y1count, y2count, z2count), decomp%even of an EVEN build (src/decomp_2d.f90:1186-1204, :448-454)"""
    out = (C.c_int64 * 4)()
    ev = C.c_int()
    lib().orc_even_counts(nx, ny, nz, p_row, p_col, rank, out, C.byref(ev))
    return tuple(out), bool(ev.value)


def best_2d_grid(nproc):
    r, c = C.c_int(), C.c_int()
    lib().orc_best_2d_grid(nproc, C.byref(r), C.byref(c))
    return r.value, c.value


def distribute(n, p):
    st, en, sz = (C.c_int * p)(), (C.c_int * p)(), (C.c_int * p)()
    lib().orc_distribute(n, p, st, en, sz)
    return list(st), list(en), list(sz)


def _ptrs(arrs):
    return (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])


def _sfx(dtype):
    dt = np.dtype(dtype)
    if dt in (np.dtype(np.float64), np.dtype(np.complex128)):
        return "_f64", np.float64, np.complex128
    if dt in (np.dtype(np.float32), np.dtype(np.complex64)):
        return "_f32", np.float32, np.complex64
    raise TypeError(dt)


_SRC_PENCIL = {X_TO_Y: 0, Y_TO_Z: 1, Z_TO_Y: 2, Y_TO_X: 1}
_DST_PENCIL = {X_TO_Y: 1, Y_TO_Z: 2, Z_TO_Y: 1, Y_TO_X: 0}


def transpose_world(direction, shape, grid, srcs):
    """transpose_{x_to_y,y_to_z,z_to_y,y_to_x} for every rank of the world; returns the dst pencils."""
    nx, ny, nz = shape
    p_row, p_col = grid
    np_ = p_row * p_col
    dt = srcs[0].dtype
    dsts = []
    for r in range(np_):
        d = Decomp(nx, ny, nz, p_row, p_col, r)
        assert srcs[r].shape == tuple(d.sz(_SRC_PENCIL[direction])), (srcs[r].shape, d.sz(_SRC_PENCIL[direction]))
        assert srcs[r].flags.f_contiguous
        dsts.append(np.zeros(d.sz(_DST_PENCIL[direction]), dtype=dt, order="F"))
    lib().orc_transpose_world(direction, nx, ny, nz, p_row, p_col, _ptrs(srcs), _ptrs(dsts), dt.itemsize)
    return dsts


def spcfft(u, isign):
    """SPCFFT of the reference (src/glassman.f90:29-67) on one complex line."""
    sfx, _, cdt = _sfx(u.dtype)
    u = np.array(u, dtype=cdt, copy=True)
    work = np.empty_like(u)
    getattr(lib(), "orc_spcfft" + sfx)(C.c_void_p(u.ctypes.data), len(u), isign, C.c_void_p(work.ctypes.data))
    return u


def c2c_1m(a, axis, isign):
    """c2c_1m_x/y/z (src/fft_generic.f90:112-208); returns a transformed copy."""
    sfx, _, cdt = _sfx(a.dtype)
    a = np.array(a, dtype=cdt, order="F", copy=True)
    getattr(lib(), "orc_c2c_1m" + sfx)(C.c_void_p(a.ctypes.data), a.shape[0], a.shape[1], a.shape[2], axis, isign)
    return a


def r2c_1m(a, axis):
    """r2c_1m_x / r2c_1m_z (src/fft_generic.f90:211-294)."""
    sfx, rdt, cdt = _sfx(a.dtype)
    a = np.asfortranarray(a, dtype=rdt)
    oshape = list(a.shape)
    oshape[axis] = a.shape[axis] // 2 + 1
    out = np.zeros(oshape, dtype=cdt, order="F")
    getattr(lib(), "orc_r2c_1m" + sfx)(C.c_void_p(a.ctypes.data), a.shape[0], a.shape[1], a.shape[2], C.c_void_p(out.ctypes.data), axis)
    return out


def c2r_1m(a, n, axis):
    """c2r_1m_x / c2r_1m_z (src/fft_generic.f90:297-384); n = real length along `axis`."""
    sfx, rdt, cdt = _sfx(a.dtype)
    a = np.asfortranarray(a, dtype=cdt)
    oshape = list(a.shape)
    oshape[axis] = n
    assert a.shape[axis] == n // 2 + 1
    out = np.zeros(oshape, dtype=rdt, order="F")
    getattr(lib(), "orc_c2r_1m" + sfx)(C.c_void_p(a.ctypes.data), C.c_void_p(out.ctypes.data), oshape[0], oshape[1], oshape[2], axis)
    return out


def sp_shape(shape, fmt):
    nx, ny, nz = shape
    return (nx // 2 + 1, ny, nz) if fmt == PHYSICAL_IN_X else (nx, ny, nz // 2 + 1)


def _skip(skip):
    if skip is None:
        return None
    return (C.c_int * 3)(*[int(bool(s)) for s in skip])


def fft_3d_c2c_world(shape, grid, fmt, isign, ins, skip=None):
    """fft_3d_c2c (src/fft_common_3d.f90:9-116) on the whole world; returns the output pencils."""
    nx, ny, nz = shape
    p_row, p_col = grid
    sfx, _, cdt = _sfx(ins[0].dtype)
    xyz = (fmt == PHYSICAL_IN_X and isign == FORWARD) or (fmt == PHYSICAL_IN_Z and isign == BACKWARD)
    outs = []
    for r in range(p_row * p_col):
        d = Decomp(nx, ny, nz, p_row, p_col, r)
        assert ins[r].shape == tuple(d.xsz if xyz else d.zsz) and ins[r].flags.f_contiguous and ins[r].dtype == cdt
        outs.append(np.zeros(d.zsz if xyz else d.xsz, dtype=cdt, order="F"))
    getattr(lib(), "orc_fft_3d_c2c" + sfx)(nx, ny, nz, p_row, p_col, fmt, isign, _skip(skip), _ptrs(ins), _ptrs(outs))
    return outs


def fft_3d_r2c_world(shape, grid, fmt, ins, skip=None):
    """fft_3d_r2c (src/fft_common_3d.f90:121-189)."""
    nx, ny, nz = shape
    p_row, p_col = grid
    sfx, rdt, cdt = _sfx(ins[0].dtype)
    sx, sy, sz = sp_shape(shape, fmt)
    outs = []
    for r in range(p_row * p_col):
        ph = Decomp(nx, ny, nz, p_row, p_col, r)
        sp = Decomp(sx, sy, sz, p_row, p_col, r)
        assert ins[r].shape == tuple(ph.xsz if fmt == PHYSICAL_IN_X else ph.zsz) and ins[r].flags.f_contiguous and ins[r].dtype == rdt
        outs.append(np.zeros(sp.zsz if fmt == PHYSICAL_IN_X else sp.xsz, dtype=cdt, order="F"))
    getattr(lib(), "orc_fft_3d_r2c" + sfx)(nx, ny, nz, p_row, p_col, fmt, _skip(skip), _ptrs(ins), _ptrs(outs))
    return outs


def fft_3d_c2r_world(shape, grid, fmt, ins, skip=None):
    """fft_3d_c2r (src/fft_common_3d.f90:194-296)."""
    nx, ny, nz = shape
    p_row, p_col = grid
    sfx, rdt, cdt = _sfx(ins[0].dtype)
    sx, sy, sz = sp_shape(shape, fmt)
    outs = []
    for r in range(p_row * p_col):
        ph = Decomp(nx, ny, nz, p_row, p_col, r)
        sp = Decomp(sx, sy, sz, p_row, p_col, r)
        assert ins[r].shape == tuple(sp.zsz if fmt == PHYSICAL_IN_X else sp.xsz) and ins[r].flags.f_contiguous and ins[r].dtype == cdt
        outs.append(np.zeros(ph.xsz if fmt == PHYSICAL_IN_X else ph.zsz, dtype=rdt, order="F"))
    getattr(lib(), "orc_fft_3d_c2r" + sfx)(nx, ny, nz, p_row, p_col, fmt, _skip(skip), _ptrs(ins), _ptrs(outs))
    return outs


# ---- helpers shared by the tests: scatter / gather a global array to / from the world ----------
def scatter(glob, grid, pencil, shape=None):
    """Cut a global (nx,ny,nz) array into the `pencil` (0 x,1 y,2 z) pieces of every rank."""
    nx, ny, nz = glob.shape if shape is None else shape
    p_row, p_col = grid
    out = []
    for r in range(p_row * p_col):
        d = Decomp(nx, ny, nz, p_row, p_col, r)
        st, sz = d.st(pencil), d.sz(pencil)
        out.append(np.asfortranarray(glob[st[0]:st[0] + sz[0], st[1]:st[1] + sz[1], st[2]:st[2] + sz[2]]))
    return out


def gather(parts, shape, grid, pencil):
    nx, ny, nz = shape
    p_row, p_col = grid
    glob = np.zeros(shape, dtype=parts[0].dtype, order="F")
    for r in range(p_row * p_col):
        d = Decomp(nx, ny, nz, p_row, p_col, r)
        st, sz = d.st(pencil), d.sz(pencil)
        glob[st[0]:st[0] + sz[0], st[1]:st[1] + sz[1], st[2]:st[2] + sz[2]] = parts[r]
    return glob


def update_halo_world(glob, grid, pencil, level, periodic=(False, False, False)):
    """update_halo + halo_exchange of the reference on a simulated world (src/halo.f90:101-198, 311-399; src/halo_common.f90;
    src/halo_exchange_{x,y,z}_body.f90), numpy restatement: per-rank pencils of `glob` (pencil = 0 X, 1 Y, 2 Z) with `level`
    ghost layers on the two decomposed axes.  Two successive exchanges -- first the axis split by dims(1), then the axis
    split by dims(2) -- whose strips span the FULL extent of the other axes, ghost layers included (MPI_TYPE_VECTOR counts
    of the bodies), so the second one carries the corners.  Neighbours: MPI_CART_SHIFT of init_neighbour (src/halo.f90:55-99);
    MPI_PROC_NULL sides stay untouched (0 here: the reference leaves them uninitialised)."""
    p_row, p_col = grid
    ins = scatter(glob, grid, pencil)
    ax1 = 1 if pencil == 0 else 0
    ax2 = 1 if pencil == 2 else 2
    h = [0, 0, 0]
    h[ax1] = h[ax2] = level
    outs = []
    for a in ins:
        o = np.zeros(tuple(a.shape[i] + 2 * h[i] for i in range(3)), dtype=a.dtype, order="F")
        o[h[0]:h[0] + a.shape[0], h[1]:h[1] + a.shape[1], h[2]:h[2] + a.shape[2]] = a
        outs.append(o)
    if level == 0:
        return outs
    for ax, npd in ((ax1, p_row), (ax2, p_col)):
        per = bool(periodic[ax])
        new = [o.copy(order="F") for o in outs]
        for r in range(p_row * p_col):
            c1, c2 = r // p_col, r % p_col
            me = c1 if ax == ax1 else c2
            im = me - 1 if me > 0 else (npd - 1 if per else None)
            ip = me + 1 if me < npd - 1 else (0 if per else None)

            def rank_of(i):
                return i * p_col + c2 if ax == ax1 else c1 * p_col + i

            def sl(lo, hi):
                s = [slice(None)] * 3
                s[ax] = slice(lo, hi)
                return tuple(s)
            n = outs[r].shape[ax]
            if im is not None:  # receive from minus: that neighbour's to-plus strip (its last `level` interior layers)
                src = outs[rank_of(im)]
                new[r][sl(0, level)] = src[sl(src.shape[ax] - 2 * level, src.shape[ax] - level)]
            if ip is not None:  # receive from plus: that neighbour's to-minus strip (its first `level` interior layers)
                src = outs[rank_of(ip)]
                new[r][sl(n - level, n)] = src[sl(level, 2 * level)]
        outs = new
    return outs
