"""CPU-side checks of the C ABI: the library loads, exports every symbol include/d2d_b200.h
declares, and its host-only decomposition arithmetic equals the oracle's (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle as orc
from util import ROOT, pkg


def test_library_exports_every_declared_symbol():
    p = pkg()
    lib = p.lib()
    hdr = open(os.path.join(ROOT, "include", "d2d_b200.h")).read()
    names = sorted(set(re.findall(r"\b(d2d_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) > 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert b"sm_100a" in lib.d2d_version()


def test_kernel_registry_covers_bench_sizes():
    p = pkg()
    lib = p.lib()
    n = lib.d2d_fft_kernel_count()
    buf = ctypes.create_string_buffer(256)
    desc = []
    for i in range(n):
        assert lib.d2d_fft_kernel_describe(i, buf, 256) == 0
        desc.append(buf.value.decode())
    for size in (64, 256, 512, 1024, 2048):
        for ty in ("f64", "f32"):
            for kind in ("line", "tile"):
                for mode in ("c2c", "r2c", "c2r"):
                    assert any(d.startswith(f"n={size} {ty} {kind} {mode}") for d in desc), (size, ty, kind, mode)


@pytest.mark.parametrize("grid", [(1, 1), (1, 2), (2, 2), (2, 4), (3, 2), (4, 2), (8, 1), (1, 8)])
@pytest.mark.parametrize("shape", [(17, 13, 11), (64, 64, 64), (1024, 1024, 513), (257, 512, 512), (1025, 2048, 2048)])
def test_decomp_matches_oracle(shape, grid):
    """d2d_decomp_* vs decomp_info_init of the reference (via the oracle): sizes, dists, counts, displacements."""
    p = pkg()
    for r in range(grid[0] * grid[1]):
        d = p.DecompInfo.for_rank(*shape, grid[0], grid[1], r)
        o = orc.Decomp(*shape, grid[0], grid[1], r)
        for name in ("xsz", "ysz", "zsz", "x1dist", "y1dist", "y2dist", "z2dist", "x1cnts", "y1cnts", "y2cnts", "z2cnts",
                     "x1disp", "y1disp", "y2disp", "z2disp"):
            assert tuple(getattr(d, name)) == tuple(getattr(o, name)), (name, r)
        for a, b in (("xst", "xst"), ("yst", "yst"), ("zst", "zst"), ("xen", "xen"), ("yen", "yen"), ("zen", "zen")):
            assert tuple(x - 1 for x in getattr(d, a)) == tuple(getattr(o, b)), (a, r)  # C ABI is 1-based
        d.finalize()


def test_best_2d_grid_matches_oracle():
    p = pkg()
    for n in (1, 2, 3, 4, 6, 8, 12, 16, 24, 36, 64):
        assert p.best_2d_grid(n) == orc.best_2d_grid(n)


def test_invalid_grid_reports_reference_error():
    """decomp_2d_init_fin.f90:43-45: min(nx,ny) >= p_row and min(ny,nz) >= p_col, status + message, no abort."""
    p = pkg()
    with pytest.raises(p.Decomp2dError) as e:
        p.DecompInfo.for_rank(4, 4, 4, 8, 1, 0)
    assert "p_row" in str(e.value)


@pytest.mark.parametrize("shape,grid", [((17, 13, 11), (2, 2)), ((1024, 1024, 513), (2, 4)), ((64, 64, 64), (2, 4)), ((34, 26, 22), (4, 2))])
def test_even_counts_match_oracle_and_formula(shape, grid):
    """x1count ... z2count of the reference's EVEN builds (src/decomp_2d.f90:1197-1203: the LAST blocks are the largest) and
    decomp%even (:448-454), against the oracle restatement and the formula evaluated with numpy"""
    import numpy as np
    import oracle as orc
    p = pkg()
    nx, ny, nz = shape
    pr, pc = grid

    def dist(n, k):
        return np.array([n // k + (1 if i >= k - n % k else 0) for i in range(k)])
    for rank in range(pr * pc):
        d = p.DecompInfo.for_rank(nx, ny, nz, pr, pc, rank)
        want, ev = orc.even_counts(nx, ny, nz, pr, pc, rank)
        assert (d.x1count, d.y1count, d.y2count, d.z2count) == want and d.even == ev
        c2 = rank % pc
        c1 = rank // pc
        x1 = dist(nx, pr)[-1] * dist(ny, pr)[-1] * dist(nz, pc)[c2]
        y2 = dist(ny, pc)[-1] * dist(nz, pc)[-1] * dist(nx, pr)[c1]
        assert want == (x1, x1, y2, y2)
        assert ev == (nx % pr == 0 and ny % pr == 0 and ny % pc == 0 and nz % pc == 0)
        # a padded buffer holds every real block
        assert max(d.x1cnts) <= d.x1count and max(d.y1cnts) <= d.y1count and max(d.y2cnts) <= d.y2count and max(d.z2cnts) <= d.z2count
        d.finalize()


def test_fortran_shim_binds_only_declared_symbols():
    """the ISO_C_BINDING interfaces of 2decomp-fft_b200/fortran/*.f90 (source only: no Fortran compiler in this image) must
    name entry points that include/d2d_b200.h declares and the library exports, and cover the calls of the hot path"""
    p = pkg()
    lib = p.lib()
    hdr = open(os.path.join(ROOT, "include", "d2d_b200.h")).read()
    declared = set(re.findall(r"\b(d2d_[a-z0-9_]+)\s*\(", hdr))
    bound = set()
    fdir = os.path.join(ROOT, "2decomp-fft_b200", "fortran")
    for f in os.listdir(fdir):
        if f.endswith(".f90"):
            bound |= set(re.findall(r'bind\(C,\s*name\s*=\s*"([a-z0-9_]+)"\)', open(os.path.join(fdir, f)).read(), flags=re.I))
    assert len(bound) >= 25
    assert not (bound - declared), sorted(bound - declared)
    assert all(hasattr(lib, n) for n in bound)
    for needed in ("d2d_ctx_create_bootstrap", "d2d_decomp_create", "d2d_transpose", "d2d_fft_plan_create", "d2d_fft_3d_r2c", "d2d_fft_3d_c2r",
                   "d2d_fft_3d_c2c", "d2d_halo_update", "d2d_last_error"):
        assert needed in bound, needed


def test_kernel_registry_covers_mixed_radix_sizes():
    """compiled plans for 3 * 2^k and 5 * 2^k (csrc/fft_kernel.cuh): every mode, both precisions"""
    p = pkg()
    lib = p.lib()
    buf = ctypes.create_string_buffer(256)
    desc = []
    for i in range(lib.d2d_fft_kernel_count()):
        assert lib.d2d_fft_kernel_describe(i, buf, 256) == 0
        desc.append(buf.value.decode())
    for size in (6, 12, 24, 48, 96, 192, 384, 768, 1536, 3072, 10, 20, 40, 80, 160, 320, 640, 1280):
        for ty in ("f64", "f32"):
            for mode in ("c2c", "r2c", "c2r"):
                assert any(d.startswith(f"n={size} {ty} tile {mode}") for d in desc), (size, ty, mode)
