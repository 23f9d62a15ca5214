"""CPU check of the arbitrary-length FFT pass arithmetic: the `__host__ __device__` butterflies of csrc/fft_any.cuh (the very
functions the CUDA kernel runs, driven by the same factorisation) against a long-double DFT, for 47 lengths in fp64 and
fp32 (tools/micro/test_any_host.cu).  Needs nvcc (host code only is executed), no GPU."""
import os
import shutil
import subprocess

import pytest

from util import ROOT


def test_any_length_pass_arithmetic_on_host(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = tmp_path / "test_any_host"
    src = os.path.join(ROOT, "tools", "micro", "test_any_host.cu")
    subprocess.run([nvcc, "-std=c++17", "-O1", "--expt-relaxed-constexpr", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(exe), src],
                   check=True, capture_output=True, timeout=600)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:]
    last = r.stdout.strip().splitlines()[-1]
    assert last.startswith("worst fp64"), last


def test_compiled_plans_pass_arithmetic_on_host(tmp_path):
    """Every compiled radix plan (powers of two, 3 * 2^k, 5 * 2^k): the PassOp / Bfly code of csrc/fft_kernel.cuh that the
    kernels run, with the threads of a line emulated one after the other, against a long-double DFT
    (tools/micro/test_plans_host.cu)."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = tmp_path / "test_plans_host"
    src = os.path.join(ROOT, "tools", "micro", "test_plans_host.cu")
    subprocess.run([nvcc, "-std=c++17", "-O1", "--expt-relaxed-constexpr", "-diag-suppress", "128,177", "-gencode", "arch=compute_100a,code=sm_100a",
                    "-o", str(exe), src], check=True, capture_output=True, timeout=900)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    assert r.stdout.strip().splitlines()[-1].startswith("worst fp64")
