"""The C ABI from a compiled host: examples/fft_r2c_z.cpp (the reference's examples/fft_physical_z/fft_r2c_z.f90 restated in
C++ against include/d2d_b200.h) must compile with nothing but the header and link against libd2dfft_b200.so.  Without a GPU
the program stops at d2d_ctx_create with the library's error message and status (the path the Fortran shim maps to
decomp_2d_abort); with a GPU it must reproduce its input within the example's own criterion."""
import os
import shutil
import subprocess

import pytest

from util import ROOT


def _build_and_run(tmp_path):
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("g++ not available")
    libdir = os.path.join(ROOT, "2decomp-fft_b200", "lib")
    if not os.path.exists(os.path.join(libdir, "libd2dfft_b200.so")):
        pytest.skip("library not built")
    exe = tmp_path / "fft_r2c_z"
    subprocess.run([cxx, "-O2", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "fft_r2c_z.cpp"), "-L", libdir, "-ld2dfft_b200", f"-Wl,-rpath,{libdir}", "-o", str(exe)],
                   check=True, capture_output=True, timeout=300)
    r = subprocess.run([str(exe), "64", "32", "128", "3"], capture_output=True, text=True, timeout=300)
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        assert r.returncode == 0 and "fft_r2c_z completed" in r.stdout, r.stdout + r.stderr
    else:
        assert r.returncode == 1 and "error" in r.stderr and "cudaSetDevice" in r.stderr, r.stdout + r.stderr


def test_c_example_builds_links_and_reports_errors(tmp_path):
    _build_and_run(tmp_path)


@pytest.mark.gpu
def test_c_example_runs_on_gpu(tmp_path):
    """the compiled C++ host on the device: d2d_dev_alloc, d2d_host_alloc_pinned, d2d_host_get_device_pointer, d2d_memcpy(_async),
    the device-pointer and the host-array 3-D entry points"""
    _build_and_run(tmp_path)
