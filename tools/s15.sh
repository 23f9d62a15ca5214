#!/bin/bash
set -x
mkdir -p gpurun_out; O=gpurun_out
date +%T
timeout 1200 python -u -m pytest tests/test_gpu_multiproc.py tests/test_gpu_fft1d.py -m gpu -q -p no:cacheprovider -k "shared_gpu or long_lines" > $O/pytest_s15.log 2>&1; tail -6 $O/pytest_s15.log
date +%T
