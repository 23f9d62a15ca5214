#!/bin/bash
set -x
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -u -m pytest tests/test_gpu_fft1d.py tests/test_gpu_fft_any.py -m gpu -q --maxfail=8 -p no:cacheprovider --timeout=300 > $O/pytest_s14.log 2>&1; tail -12 $O/pytest_s14.log
