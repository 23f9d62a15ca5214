#!/bin/bash
# last sanity pass on the library rebuilt from a clean tree: the quick parts of the GPU suite within the remaining budget
set -x
mkdir -p gpurun_out; O=gpurun_out
timeout 170 python -u -m pytest tests/test_gpu_fft3d.py tests/test_gpu_configs.py tests/test_gpu_fft1d.py tests/test_c_example.py -m gpu -q -x -p no:cacheprovider --timeout=120 > $O/pytest_s27.log 2>&1; tail -4 $O/pytest_s27.log
