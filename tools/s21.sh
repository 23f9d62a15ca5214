#!/bin/bash
set -x
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -u -m pytest tests/test_gpu_fft_any.py tests/test_gpu_multiple_grids.py -m gpu -q --maxfail=20 -p no:cacheprovider --timeout=300 > $O/pytest_s21.log 2>&1; tail -5 $O/pytest_s21.log
timeout 600 python -u -m pytest tests/test_gpu_fft1d.py -m gpu -q --maxfail=20 -p no:cacheprovider --timeout=300 -k 'chirp or large_prime or unsupported or long_lines' > $O/pytest_s21b.log 2>&1; tail -3 $O/pytest_s21b.log
run() { echo "== $*"; env "$@" timeout 200 python tools/bench_kernels.py --shape $SHAPE --prec ${PREC:-f64} --reps 3 --only3d --fmt Z 2>&1 | tail -7; }
SHAPE=510,510,510 run D2D_ANY_BIG=0
SHAPE=510,510,510 run D2D_ANY_BIG=1
SHAPE=510,510,510 run D2D_ANY_BIG=1 D2D_ANY_ASYNC=0
SHAPE=360,360,360 run D2D_ANY_BIG=0
SHAPE=360,360,360 run D2D_ANY_BIG=1
SHAPE=544,416,352 run D2D_ANY_BIG=0
SHAPE=544,416,352 run D2D_ANY_BIG=1
SHAPE=1000,1000,1000 run D2D_ANY_BIG=1
SHAPE=1000,1000,1000 run D2D_ANY_BIG=2
PREC=f32 SHAPE=510,510,510 run D2D_ANY_BIG=1
PREC=f32 SHAPE=510,510,510 run D2D_ANY_BIG=0
