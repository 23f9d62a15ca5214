"""Host-side mirror of the 2DECOMP&FFT API for the hot path, on top of the C ABI (include/d2d_b200.h).

The reference is Fortran; this image has no Fortran compiler, so the drop-in Fortran shim ships as
source (2decomp-fft_b200/fortran/) and this module is the executable mirror of the same interface:
same names, argument meaning and error behaviour as

    decomp_2d_init / decomp_2d_finalize            src/decomp_2d_init_fin.f90:15-226
    decomp_info_init / decomp_info                  src/decomp_2d.f90:382-515, src/info.f90:19-47
    transpose_x_to_y / y_to_z / z_to_y / y_to_x     src/transpose_*.f90
    alloc_x / alloc_y / alloc_z                     src/alloc.f90:10-134 (device: src/alloc_dev.f90)
    decomp_2d_fft_init / _3d / _finalize / _get_size  src/fft_common.f90:42-327, src/fft_cufft.f90:676-1170

PyTorch is plumbing only (device memory + torch.distributed for the unique-id broadcast); every
compute call goes through ctypes into libd2dfft_b200.so.  There is NO fallback: if the shared
library is missing, importing any compute entry point raises.

Arrays are torch CUDA tensors viewed in Fortran order: shape (n1, n2, n3) with strides (1, n1, n1*n2),
as returned by alloc_x/alloc_y/alloc_z.

Module-level state ("current decomposition", "current FFT engine") is thread-local, so that several
ranks can live in one process as threads (d2d_group transport) -- the reference keeps the same state
in module variables of one MPI process.
"""
import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libd2dfft_b200.so")
if os.environ.get("D2D_B200_LIB"):  # experiments: an alternative build of the same library (Makefile VARIANT=...)
    LIB_PATH = os.environ["D2D_B200_LIB"]

# src/decomp_2d_constants.f90:86-93
DECOMP_2D_FFT_FORWARD = -1
DECOMP_2D_FFT_BACKWARD = 1
PHYSICAL_IN_X = 1
PHYSICAL_IN_Z = 3
D2D_F32, D2D_F64 = 0, 1
X_TO_Y, Y_TO_Z, Z_TO_Y, Y_TO_X = 0, 1, 2, 3


class Decomp2dError(RuntimeError):
    """What decomp_2d_abort (src/decomp_2d_mpi.f90:145-191) reports: error code + message."""

    def __init__(self, code, msg):
        super().__init__(f"2DECOMP&FFT ERROR - errorcode: {code}\nERROR MESSAGE: {msg}")
        self.errorcode = code
        self.msg = msg


_lib = None


def lib():
    """Load libd2dfft_b200.so (fails loudly when it was not built: there is no CPU/PyTorch fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no fallback path)")
        try:
            # torch bundles a newer NCCL (libnccl.so.2, same soname as the system one this library links to):
            # whichever is loaded first serves both, so let torch's superset win
            import torch  # noqa: F401
        except ImportError:
            pass
        l = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        l.d2d_last_error.restype = C.c_char_p
        l.d2d_version.restype = C.c_char_p
        l.d2d_ctx_stream.restype = C.c_void_p
        l.d2d_ctx_launch_count.restype = C.c_int64
        for name in ("d2d_transpose", ):
            getattr(l, name).argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        for name in ("d2d_transpose_x_to_y", "d2d_transpose_y_to_z", "d2d_transpose_z_to_y", "d2d_transpose_y_to_x"):
            getattr(l, name).argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        l.d2d_fft_3d_c2c.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        l.d2d_fft_3d_r2c.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        l.d2d_fft_3d_c2r.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        l.d2d_fft_3d_c2c_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        l.d2d_fft_3d_r2c_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        l.d2d_fft_3d_c2r_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        l.d2d_fft_c2c_1m.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        l.d2d_fft_r2c_1m.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        l.d2d_fft_c2r_1m.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        l.d2d_ctx_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        l.d2d_ctx_create_bootstrap.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        l.d2d_ctx_create_in_group.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        l.d2d_ctx_destroy.argtypes = [C.c_void_p]
        l.d2d_ctx_sync.argtypes = [C.c_void_p]
        l.d2d_ctx_set_blocking.argtypes = [C.c_void_p, C.c_int]
        l.d2d_ctx_stream.argtypes = [C.c_void_p]
        l.d2d_ctx_launch_count.argtypes = [C.c_void_p]
        l.d2d_ctx_profile.argtypes = [C.c_void_p, C.c_int]
        l.d2d_ctx_profile_count.argtypes = [C.c_void_p]
        l.d2d_ctx_profile_reset.argtypes = [C.c_void_p]
        l.d2d_ctx_profile_get.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p]
        l.d2d_ctx_info.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        l.d2d_group_create.argtypes = [C.c_void_p, C.c_int]
        l.d2d_group_destroy.argtypes = [C.c_void_p]
        l.d2d_group_abort.argtypes = [C.c_void_p]
        l.d2d_decomp_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        l.d2d_decomp_create_for_rank.argtypes = [C.c_int] * 6 + [C.c_void_p]
        l.d2d_decomp_destroy.argtypes = [C.c_void_p]
        l.d2d_decomp_query.argtypes = [C.c_void_p] * 10
        l.d2d_decomp_dist.argtypes = [C.c_void_p] * 5
        l.d2d_decomp_counts.argtypes = [C.c_void_p] * 9
        l.d2d_decomp_even.argtypes = [C.c_void_p] * 6
        l.d2d_halo_update.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        l.d2d_ctx_set_even.argtypes = [C.c_void_p, C.c_int]
        l.d2d_fft_plan_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        l.d2d_fft_plan_destroy.argtypes = [C.c_void_p]
        l.d2d_fft_plan_ph.argtypes = [C.c_void_p, C.c_void_p]
        l.d2d_fft_plan_sp.argtypes = [C.c_void_p, C.c_void_p]
        l.d2d_fft_get_size.argtypes = [C.c_void_p] * 4
        l.d2d_host_alloc_pinned.argtypes = [C.c_void_p, C.c_int64]
        l.d2d_host_free.argtypes = [C.c_void_p]
        _lib = l
    return _lib


def _check(status):
    if status != 0:
        raise Decomp2dError(status, lib().d2d_last_error().decode())


def get_unique_id():
    """ncclGetUniqueId on rank 0 (src/decomp_2d_nccl.f90:182-185); broadcast it yourself."""
    buf = (C.c_ubyte * 128)()
    _check(lib().d2d_get_unique_id(buf))
    return bytes(buf)


def best_2d_grid(nproc):
    r, c = C.c_int(), C.c_int()
    _check(lib().d2d_best_2d_grid(nproc, C.byref(r), C.byref(c)))
    return r.value, c.value


class DecompInfo:
    """decomp_info (src/info.f90:19-47): xst/xen/xsz ... (1-based starts like the reference)."""

    def __init__(self, handle, p_row, p_col, owned=True):
        self._h = handle
        self._owned = owned
        l = lib()
        a = [(C.c_int * 3)() for _ in range(9)]
        _check(l.d2d_decomp_query(handle, *a))
        (self.xst, self.xen, self.xsz, self.yst, self.yen, self.ysz, self.zst, self.zen, self.zsz) = [tuple(x) for x in a]
        d = [(C.c_int * p_row)(), (C.c_int * p_row)(), (C.c_int * p_col)(), (C.c_int * p_col)()]
        _check(l.d2d_decomp_dist(handle, *d))
        self.x1dist, self.y1dist, self.y2dist, self.z2dist = [tuple(x) for x in d]
        c = [(C.c_int64 * p_row)(), (C.c_int64 * p_row)(), (C.c_int64 * p_col)(), (C.c_int64 * p_col)(),
             (C.c_int64 * p_row)(), (C.c_int64 * p_row)(), (C.c_int64 * p_col)(), (C.c_int64 * p_col)()]
        _check(l.d2d_decomp_counts(handle, *c))
        (self.x1cnts, self.y1cnts, self.y2cnts, self.z2cnts, self.x1disp, self.y1disp, self.y2disp, self.z2disp) = [tuple(x) for x in c]
        e = [C.c_int64() for _ in range(4)]
        ev = C.c_int()
        _check(l.d2d_decomp_even(handle, *[C.byref(v) for v in e], C.byref(ev)))
        self.x1count, self.y1count, self.y2count, self.z2count = [v.value for v in e]  # EVEN builds (decomp_2d.f90:1197-1203)
        self.even = bool(ev.value)

    @classmethod
    def for_rank(cls, nx, ny, nz, p_row, p_col, rank):
        """Host-only decomposition record of any rank (no device needed)."""
        h = C.c_void_p()
        _check(lib().d2d_decomp_create_for_rank(nx, ny, nz, p_row, p_col, rank, C.byref(h)))
        return cls(h, p_row, p_col)

    def finalize(self):
        if self._h is not None and self._owned:
            lib().d2d_decomp_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.finalize()
        except Exception:
            pass


class Group:
    """In-process rank group (thread-per-rank transport, d2d_group_*)."""

    def __init__(self, nranks):
        self._h = C.c_void_p()
        self.nranks = nranks
        _check(lib().d2d_group_create(C.byref(self._h), nranks))

    def abort(self):
        """A rank failed: wake the ranks waiting for it in an exchange (they raise Decomp2dError)."""
        if self._h:
            lib().d2d_group_abort(self._h)

    def destroy(self):
        if self._h:
            lib().d2d_group_destroy(self._h)
            self._h = None


def _torch():
    import torch
    return torch


def _dtype_code(t):
    torch = _torch()
    if t.dtype in (torch.float64, torch.complex128):
        return D2D_F64
    if t.dtype in (torch.float32, torch.complex64):
        return D2D_F32
    raise Decomp2dError(2, f"unsupported dtype {t.dtype}")


_ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64)


def _allgather_callback(fn):
    """C callback (d2d_allgather_fn) around fn(bytes) -> [bytes of rank 0, bytes of rank 1, ...]."""
    def cb(_user, send, recv, nbytes):
        try:
            parts = fn(C.string_at(send, nbytes))
            blob = b"".join(parts)
            C.memmove(recv, blob, len(blob))
            return 0
        except Exception:  # the C side turns this into a Decomp2dError
            import traceback
            traceback.print_exc()
            return 1
    return _ALLGATHER_FN(cb)


def _check_dtype(t, engine_dtype, what):
    torch = _torch()
    want = {torch.float64: (torch.float64, torch.complex128), torch.float32: (torch.float32, torch.complex64)}[engine_dtype]
    if t.dtype not in want:
        raise Decomp2dError(2, f"{what}: dtype {t.dtype} does not match the engine precision {engine_dtype}")


def _check_pencil(t, shape, what):
    if tuple(t.shape) != tuple(shape):
        raise Decomp2dError(2, f"{what}: array shape {tuple(t.shape)} does not match the pencil {tuple(shape)}")
    n1, n2, _ = shape
    if t.numel() and tuple(t.stride()) != (1, n1, n1 * n2):
        raise Decomp2dError(2, f"{what}: array must be Fortran-ordered (use alloc_x/alloc_y/alloc_z)")
    if not t.is_cuda:
        raise Decomp2dError(2, f"{what}: array must live on the device")


class Decomp2d:
    """One rank's library state: what decomp_2d_init sets up (src/decomp_2d_init_fin.f90:15-184)."""

    def __init__(self, nx, ny, nz, p_row, p_col, rank=0, nranks=1, unique_id=None, device=None, group=None, allgather=None,
                 periodic_bc=None):
        """unique_id: NCCL bootstrap (d2d_ctx_create); group: thread-per-rank ranks of one process; allgather: a callable
        bytes -> [bytes per rank] (MPI_Allgather-like) for the NCCL-free bootstrap (d2d_ctx_create_bootstrap)."""
        torch = _torch()
        l = lib()
        if p_row <= 0 or p_col <= 0:  # auto-tuning mode (decomp_2d_init_fin.f90:55-77)
            p_row, p_col = best_2d_grid(nranks)
        if device is None:
            device = torch.cuda.current_device()
        self.device = device
        self._h = C.c_void_p()
        self._allgather_cb = None
        # periodic_bc of decomp_2d_init (src/decomp_2d_init_fin.f90:31-41): only the halo exchange looks at it
        self.periodic_bc = tuple(bool(v) for v in (periodic_bc if periodic_bc is not None else (False, False, False)))
        if group is not None:
            _check(l.d2d_ctx_create_in_group(C.byref(self._h), group._h, rank, p_row, p_col, device))
        elif allgather is not None:
            self._allgather_cb = _allgather_callback(allgather)  # must outlive the context
            _check(l.d2d_ctx_create_bootstrap(C.byref(self._h), nranks, rank, p_row, p_col, device,
                                              C.cast(self._allgather_cb, C.c_void_p), None))
        else:
            idbuf = (C.c_ubyte * 128).from_buffer_copy(unique_id) if unique_id is not None else None
            _check(l.d2d_ctx_create(C.byref(self._h), idbuf, nranks, rank, p_row, p_col, device))
        self.nrank, self.nproc = rank, nranks
        self.dims = (p_row, p_col)
        self.coord = (rank // p_col, rank % p_col)
        self.nx_global, self.ny_global, self.nz_global = nx, ny, nz
        self.decomp_main = self.decomp_info_init(nx, ny, nz)
        d = self.decomp_main
        self.xstart, self.xend, self.xsize = d.xst, d.xen, d.xsz
        self.ystart, self.yend, self.ysize = d.yst, d.yen, d.ysz
        self.zstart, self.zend, self.zsize = d.zst, d.zen, d.zsz

    # decomp_info_init (src/decomp_2d.f90:382-490)
    def decomp_info_init(self, nx, ny, nz):
        h = C.c_void_p()
        _check(lib().d2d_decomp_create(self._h, nx, ny, nz, C.byref(h)))
        return DecompInfo(h, *self.dims)

    # alloc_x / alloc_y / alloc_z (src/alloc.f90:10-134; device arrays: src/alloc_dev.f90): pencil-shaped array, Fortran order.
    # opt_levels = halo cells added on both sides of each dimension (alloc.f90:39-59 via opt_levels); opt_global = the
    # Fortran array would be indexed with global bounds (xstart:xend): recorded as the `lbound` attribute of the tensor, whose
    # element [0,0,0] is then global point lbound (1-based, like the reference), otherwise lbound = (1,1,1) - levels.
    def _alloc(self, pencil, dtype, decomp, opt_global=False, opt_levels=None):
        torch = _torch()
        d = decomp or self.decomp_main
        lv = tuple(int(v) for v in (opt_levels if opt_levels is not None else (0, 0, 0)))
        if len(lv) != 3 or min(lv) < 0:
            raise Decomp2dError(2, "opt_levels must be three non-negative integers")
        sz = (d.xsz, d.ysz, d.zsz)[pencil]
        st = (d.xst, d.yst, d.zst)[pencil]
        n1, n2, n3 = [sz[i] + 2 * lv[i] for i in range(3)]
        base = torch.zeros((n3, n2, n1), dtype=dtype, device=f"cuda:{self.device}")
        t = base.permute(2, 1, 0)
        t.lbound = tuple((st[i] if opt_global else 1) - lv[i] for i in range(3))
        return t

    def alloc_x(self, dtype, decomp=None, opt_global=False, opt_levels=None):
        return self._alloc(0, dtype, decomp, opt_global, opt_levels)

    def alloc_y(self, dtype, decomp=None, opt_global=False, opt_levels=None):
        return self._alloc(1, dtype, decomp, opt_global, opt_levels)

    def alloc_z(self, dtype, decomp=None, opt_global=False, opt_levels=None):
        return self._alloc(2, dtype, decomp, opt_global, opt_levels)

    def _transpose(self, direction, src, dst, decomp):
        d = decomp or self.decomp_main
        sizes = (d.xsz, d.ysz, d.zsz)
        _check_pencil(src, sizes[(0, 1, 2, 1)[direction]], "src")
        _check_pencil(dst, sizes[(1, 2, 1, 0)[direction]], "dst")
        if src.dtype != dst.dtype:
            raise Decomp2dError(2, "src and dst must have the same type")
        _check(lib().d2d_transpose(self._h, d._h, direction, _dtype_code(src), int(src.is_complex()), src.data_ptr(), dst.data_ptr()))

    # update_halo (src/halo.f90:101-198): returns `out`, the pencil with `level` ghost layers on its two decomposed axes,
    # interior copied from `in` and ghost layers filled from the neighbouring pencils.  opt_pencil = 1, 2, 3 like the reference
    # (deduced from the shape when omitted, src/halo.f90:201-245); opt_global only changes the Fortran index bounds of `out`
    # (recorded as `lbound`).
    def update_halo(self, inp, level, decomp=None, opt_global=False, opt_pencil=None):
        torch = _torch()
        d = decomp or self.decomp_main
        shp = tuple(inp.shape)
        if opt_pencil is None:
            if shp[0] == d.xsz[0]:
                opt_pencil = 1
            elif shp[1] == d.ysz[1]:
                opt_pencil = 2
            elif shp[2] == d.zsz[2]:
                opt_pencil = 3
            else:
                raise Decomp2dError(1, "Invalid decomposition size")
        if opt_pencil not in (1, 2, 3):
            raise Decomp2dError(10, "Invalid data passed to update_halo")
        pen = opt_pencil - 1
        _check_pencil(inp, (d.xsz, d.ysz, d.zsz)[pen], "in")
        lv = [level, level, level]
        lv[pen] = 0
        out = self._alloc(pen, inp.dtype, d, opt_global, lv)
        per = (C.c_int * 3)(*[int(v) for v in self.periodic_bc])
        _check(lib().d2d_halo_update(self._h, d._h, pen, level, _dtype_code(inp), int(inp.is_complex()), per, inp.data_ptr(), out.data_ptr()))
        return out

    def transpose_x_to_y(self, src, dst, decomp=None):
        self._transpose(X_TO_Y, src, dst, decomp)

    def transpose_y_to_z(self, src, dst, decomp=None):
        self._transpose(Y_TO_Z, src, dst, decomp)

    def transpose_z_to_y(self, src, dst, decomp=None):
        self._transpose(Z_TO_Y, src, dst, decomp)

    def transpose_y_to_x(self, src, dst, decomp=None):
        self._transpose(Y_TO_X, src, dst, decomp)

    # bare batched 1-D transforms on a local array (c2c_1m_x/y/z ..., src/fft_cufft.f90:489-671)
    @staticmethod
    def _check_1m(a, out, shape_a, shape_out):
        for t, shp, what in ((a, shape_a, "in"), (out, shape_out, "out")):
            _check_pencil(t, shp, what)
        if _dtype_code(a) != _dtype_code(out):
            raise Decomp2dError(2, "in and out must have the same precision")

    def c2c_1m(self, a, axis, isign, out=None):
        out = a if out is None else out
        n1, n2, n3 = a.shape
        self._check_1m(a, out, a.shape, a.shape)
        if not (a.is_complex() and out.is_complex()):
            raise Decomp2dError(2, "c2c_1m works on complex arrays")
        _check(lib().d2d_fft_c2c_1m(self._h, _dtype_code(a), axis, n1, n2, n3, a.data_ptr(), out.data_ptr(), isign))
        return out

    def r2c_1m(self, a, out, axis):
        n1, n2, n3 = a.shape
        cs = [n1, n2, n3]
        cs[axis] = cs[axis] // 2 + 1
        self._check_1m(a, out, a.shape, cs)
        if a.is_complex() or not out.is_complex():
            raise Decomp2dError(2, "r2c_1m: real input, complex output")
        _check(lib().d2d_fft_r2c_1m(self._h, _dtype_code(a), axis, n1, n2, n3, a.data_ptr(), out.data_ptr()))
        return out

    def c2r_1m(self, a, out, axis):
        n1, n2, n3 = out.shape
        cs = [n1, n2, n3]
        cs[axis] = cs[axis] // 2 + 1
        self._check_1m(a, out, cs, out.shape)
        if not a.is_complex() or out.is_complex():
            raise Decomp2dError(2, "c2r_1m: complex input, real output")
        _check(lib().d2d_fft_c2r_1m(self._h, _dtype_code(a), axis, n1, n2, n3, a.data_ptr(), out.data_ptr()))
        return out

    def sync(self):
        _check(lib().d2d_ctx_sync(self._h))

    def set_blocking(self, flag):
        _check(lib().d2d_ctx_set_blocking(self._h, int(flag)))

    def set_even(self, flag):
        """bare transposes in the padded equal-count layout of the reference's EVEN builds (same pencils)"""
        _check(lib().d2d_ctx_set_even(self._h, int(flag)))

    def stream(self):
        return lib().d2d_ctx_stream(self._h)

    def launch_count(self):
        return lib().d2d_ctx_launch_count(self._h)

    def profile(self, enable):
        _check(lib().d2d_ctx_profile(self._h, int(enable)))

    def profile_reset(self):
        _check(lib().d2d_ctx_profile_reset(self._h))

    def profile_read(self):
        """{label: (total_ms, calls, bytes)} of the per-stage device timers."""
        l = lib()
        out = {}
        for i in range(l.d2d_ctx_profile_count(self._h)):
            label = C.create_string_buffer(64)
            ms, calls, by = C.c_double(), C.c_int64(), C.c_double()
            _check(l.d2d_ctx_profile_get(self._h, i, label, C.byref(ms), C.byref(calls), C.byref(by)))
            out[label.value.decode()] = (ms.value, calls.value, by.value)
        return out

    # decomp_2d_finalize (src/decomp_2d_init_fin.f90:189-226)
    def finalize(self):
        if self._h:
            if self.decomp_main is not None:
                self.decomp_main.finalize()
                self.decomp_main = None
            lib().d2d_ctx_destroy(self._h)
            self._h = None


class Decomp2dFFTEngine:
    """decomp_2d_fft_engine (src/fft_cufft.f90:37-61; init: src/fft_common.f90:141-242)."""

    def __init__(self, d2d, pencil, nx=None, ny=None, nz=None, dtype=None, opt_inplace=False, opt_inplace_r2c=False,
                 opt_inplace_c2r=False, opt_skip_XYZ_c2c=None):
        torch = _torch()
        if pencil not in (PHYSICAL_IN_X, PHYSICAL_IN_Z):
            raise Decomp2dError(1, "Invalid value for format")
        if opt_inplace_r2c or opt_inplace_c2r:
            # src/fft_common.f90:177-182: only the fftw_f03 backend supports in-place r2c / c2r
            raise Decomp2dError(1, "In-place r2c / c2r transforms are not supported by this backend")
        self.d2d = d2d
        self.format = pencil
        self.nx_fft = d2d.nx_global if nx is None else nx
        self.ny_fft = d2d.ny_global if ny is None else ny
        self.nz_fft = d2d.nz_global if nz is None else nz
        self.dtype = torch.float64 if dtype is None else dtype
        self.inplace = bool(opt_inplace)
        self.inplace_r2c = self.inplace_c2r = False  # never supported by this backend (checked above), like cuFFT's
        skip = list(opt_skip_XYZ_c2c) if opt_skip_XYZ_c2c is not None else [False] * 3
        self.skip_x_c2c, self.skip_y_c2c, self.skip_z_c2c = [bool(s) for s in skip]
        cskip = (C.c_int * 3)(*[int(bool(s)) for s in skip])
        self._h = C.c_void_p()
        code = D2D_F64 if self.dtype == torch.float64 else D2D_F32
        _check(lib().d2d_fft_plan_create(d2d._h, pencil, self.nx_fft, self.ny_fft, self.nz_fft, code, int(self.inplace), cskip,
                                         C.byref(self._h)))
        h = C.c_void_p()
        _check(lib().d2d_fft_plan_ph(self._h, C.byref(h)))
        self.ph = DecompInfo(h, *d2d.dims, owned=False)
        h = C.c_void_p()
        _check(lib().d2d_fft_plan_sp(self._h, C.byref(h)))
        self.sp = DecompInfo(h, *d2d.dims, owned=False)

    @property
    def real_dtype(self):
        return self.dtype

    @property
    def complex_dtype(self):
        torch = _torch()
        return torch.complex128 if self.dtype == torch.float64 else torch.complex64

    # decomp_2d_fft_get_size (src/fft_common.f90:311-327)
    def get_size(self):
        a = [(C.c_int * 3)() for _ in range(3)]
        _check(lib().d2d_fft_get_size(self._h, *a))
        return tuple(a[0]), tuple(a[1]), tuple(a[2])

    def _pencils(self, isign=None):
        fx = self.format == PHYSICAL_IN_X
        return fx

    # decomp_2d_fft_3d generic interface (src/fft_common.f90:31-38): (in_c,out_c,isign) | (in_r,out_c) | (in_c,out_r)
    def fft_3d(self, inp, out, isign=None):
        fx = self.format == PHYSICAL_IN_X
        _check_dtype(inp, self.dtype, "in")
        _check_dtype(out, self.dtype, "out")
        if inp.data_ptr() == out.data_ptr() and inp.numel():
            raise Decomp2dError(2, "decomp_2d_fft_3d: the input and output arrays overlap")
        if inp.is_complex() and out.is_complex():
            if isign not in (DECOMP_2D_FFT_FORWARD, DECOMP_2D_FFT_BACKWARD):
                raise Decomp2dError(1, "c2c transforms need isign = DECOMP_2D_FFT_FORWARD / _BACKWARD")
            xyz = (fx and isign == DECOMP_2D_FFT_FORWARD) or ((not fx) and isign == DECOMP_2D_FFT_BACKWARD)
            _check_pencil(inp, self.ph.xsz if xyz else self.ph.zsz, "in")
            _check_pencil(out, self.ph.zsz if xyz else self.ph.xsz, "out")
            _check(lib().d2d_fft_3d_c2c(self._h, inp.data_ptr(), out.data_ptr(), isign))
        elif (not inp.is_complex()) and out.is_complex():
            _check_pencil(inp, self.ph.xsz if fx else self.ph.zsz, "in_r")
            _check_pencil(out, self.sp.zsz if fx else self.sp.xsz, "out_c")
            _check(lib().d2d_fft_3d_r2c(self._h, inp.data_ptr(), out.data_ptr()))
        elif inp.is_complex() and not out.is_complex():
            _check_pencil(inp, self.sp.zsz if fx else self.sp.xsz, "in_c")
            _check_pencil(out, self.ph.xsz if fx else self.ph.zsz, "out_r")
            _check(lib().d2d_fft_3d_c2r(self._h, inp.data_ptr(), out.data_ptr()))
        else:
            raise Decomp2dError(1, "real to real transforms are not part of decomp_2d_fft_3d")

    # host-array variants (what `!$acc data copyin(in) copy(out)` does around the reference's calls)
    def fft_3d_r2c_host(self, in_host_ptr, out_host_ptr):
        _check(lib().d2d_fft_3d_r2c_host(self._h, in_host_ptr, out_host_ptr))

    def fft_3d_c2r_host(self, in_host_ptr, out_host_ptr):
        _check(lib().d2d_fft_3d_c2r_host(self._h, in_host_ptr, out_host_ptr))

    def fft_3d_c2c_host(self, in_host_ptr, out_host_ptr, isign):
        _check(lib().d2d_fft_3d_c2c_host(self._h, in_host_ptr, out_host_ptr, isign))

    def fin(self):
        if self._h:
            lib().d2d_fft_plan_destroy(self._h)
            self._h = None


# ---- module-level API with the reference's names (thread-local "module variables") ----------------
_state = threading.local()


def decomp_2d_init(nx, ny, nz, p_row, p_col, rank=0, nranks=1, unique_id=None, device=None, group=None, allgather=None,
                   periodic_bc=None):
    _state.d2d = Decomp2d(nx, ny, nz, p_row, p_col, rank=rank, nranks=nranks, unique_id=unique_id, device=device, group=group,
                          allgather=allgather, periodic_bc=periodic_bc)
    _state.engines = {}
    _state.current = None
    _state.n_grid = 0
    return _state.d2d


def decomp_2d_init_from_torch_distributed(nx, ny, nz, p_row, p_col, transport=None):
    """decomp_2d_init for a torchrun job.  transport "nccl" (default on an NCCL process group): the unique id travels
    through torch.distributed, standing in for the MPI_Bcast of src/decomp_2d_nccl.f90:185, and NCCL stays available as the
    second data plane (D2D_P2P=0).  transport "boot" (default on any other process group, or D2D_TRANSPORT=boot): no NCCL at
    all -- torch.distributed only all-gathers the CUDA-IPC handles at plan creation (what MPI_Allgather does in the Fortran
    shim) and the exchange is the library's own peer-memory path."""
    torch = _torch()
    import torch.distributed as dist
    rank, nranks = dist.get_rank(), dist.get_world_size()
    if transport is None:
        transport = os.environ.get("D2D_TRANSPORT") or ("nccl" if dist.get_backend() == "nccl" else "boot")
    if transport == "boot" and nranks > 1:
        def allgather(data):
            parts = [None] * nranks
            dist.all_gather_object(parts, data)
            return parts
        return decomp_2d_init(nx, ny, nz, p_row, p_col, rank=rank, nranks=nranks, device=torch.cuda.current_device(),
                              allgather=allgather)
    obj = [get_unique_id() if rank == 0 else None]
    if nranks > 1:
        dist.broadcast_object_list(obj, src=0)
    return decomp_2d_init(nx, ny, nz, p_row, p_col, rank=rank, nranks=nranks, unique_id=obj[0], device=torch.cuda.current_device())


def _cur():
    d = getattr(_state, "d2d", None)
    if d is None:
        raise Decomp2dError(1, "decomp_2d_init has not been called")
    return d


def decomp_2d_finalize():
    decomp_2d_fft_finalize()
    _cur().finalize()
    _state.d2d = None


def get_decomp_info():
    return _cur().decomp_main


def update_halo(inp, level, decomp=None, opt_global=False, opt_pencil=None):
    return _cur().update_halo(inp, level, decomp, opt_global, opt_pencil)


def transpose_x_to_y(src, dst, decomp=None):
    _cur().transpose_x_to_y(src, dst, decomp)


def transpose_y_to_z(src, dst, decomp=None):
    _cur().transpose_y_to_z(src, dst, decomp)


def transpose_z_to_y(src, dst, decomp=None):
    _cur().transpose_z_to_y(src, dst, decomp)


def transpose_y_to_x(src, dst, decomp=None):
    _cur().transpose_y_to_x(src, dst, decomp)


def alloc_x(dtype, decomp=None, opt_global=False, opt_levels=None):
    return _cur().alloc_x(dtype, decomp, opt_global, opt_levels)


def alloc_y(dtype, decomp=None, opt_global=False, opt_levels=None):
    return _cur().alloc_y(dtype, decomp, opt_global, opt_levels)


def alloc_z(dtype, decomp=None, opt_global=False, opt_levels=None):
    return _cur().alloc_z(dtype, decomp, opt_global, opt_levels)


# decomp_2d_fft_set_ngrid / _get_ngrid (src/fft_common.f90:368-413): the number of FFT engines the module keeps.  Engines
# that exist survive a resize (up to the new count), like the move_alloc of the reference.
def decomp_2d_fft_set_ngrid(ngrd):
    _cur()
    if ngrd < 1:
        raise Decomp2dError(ngrd, "Invalid value for n_grid")
    engines = getattr(_state, "engines", {})
    for ig in [k for k in engines if k > ngrd]:
        engines.pop(ig).fin()
    _state.engines = engines
    _state.n_grid = ngrd


def decomp_2d_fft_get_ngrid():
    return getattr(_state, "n_grid", 0)


def decomp_2d_fft_init(pencil=PHYSICAL_IN_X, nx=None, ny=None, nz=None, igrid=None, **opts):
    """decomp_2d_fft_init, all four overloads (src/fft_common.f90:42-138): without igrid ONE engine is kept
    (fft_init_one_grid: set_ngrid(1), engine 1); with igrid the engine of that grid is (re)initialised
    (fft_init_multiple_grids; decomp_2d_fft_set_ngrid must have been called)."""
    if igrid is None:
        decomp_2d_fft_set_ngrid(1)
        igrid = 1
    elif igrid < 1 or igrid > decomp_2d_fft_get_ngrid():
        raise Decomp2dError(igrid, "Invalid value for igrid")
    eng = Decomp2dFFTEngine(_cur(), pencil, nx, ny, nz, **opts)
    old = _state.engines.get(igrid)
    if old is not None:
        old.fin()
    _state.engines[igrid] = eng
    _state.current = eng
    return eng


def _grid_engine(igrid):
    n = decomp_2d_fft_get_ngrid()
    if n < 1:
        raise Decomp2dError(n, "The FFT module was not initialised")
    if igrid < 1 or igrid > n:
        raise Decomp2dError(igrid, "Invalid value for igrid")
    eng = _state.engines.get(igrid)
    if eng is None or not eng._h:
        raise Decomp2dError(-1, "FFT engine is not ready")
    return eng


# decomp_2d_fft_use_grid (src/fft_common.f90:418-443) -> engine%use_it (:446-479)
def decomp_2d_fft_use_grid(igrid=1):
    _state.current = _grid_engine(igrid)
    return _state.current


# decomp_2d_fft_get_engine (src/fft_common.f90:485-511)
def decomp_2d_fft_get_engine(igrid=1):
    return _grid_engine(igrid)


def _current_engine():
    eng = getattr(_state, "current", None)
    if eng is None:
        raise Decomp2dError(1, "decomp_2d_fft_init has not been called")
    return eng


def decomp_2d_fft_3d(inp, out, isign=None):
    _current_engine().fft_3d(inp, out, isign)


def decomp_2d_fft_get_size():
    return _current_engine().get_size()


def decomp_2d_fft_get_ph():
    return _current_engine().ph


def decomp_2d_fft_get_sp():
    return _current_engine().sp


def decomp_2d_fft_get_format():
    return _current_engine().format


# src/fft_common.f90:525-555
def decomp_2d_fft_get_inplace():
    return _current_engine().inplace


def decomp_2d_fft_get_inplace_r2c():
    return _current_engine().inplace_r2c


def decomp_2d_fft_get_inplace_c2r():
    return _current_engine().inplace_c2r


def decomp_2d_fft_finalize():
    for eng in getattr(_state, "engines", {}).values():
        eng.fin()
    _state.engines = {}
    _state.current = None
    _state.n_grid = 0
