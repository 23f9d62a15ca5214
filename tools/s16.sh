#!/bin/bash
set -x
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -u -m pytest tests/test_gpu_host_arrays.py -m gpu -q --maxfail=8 -p no:cacheprovider --timeout=300 > $O/pytest_s16.log 2>&1; tail -15 $O/pytest_s16.log
