#!/bin/bash
set -x
mkdir -p gpurun_out; O=gpurun_out
timeout 1500 python -u -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider --timeout=600 > $O/pytest_s22.log 2>&1; tail -8 $O/pytest_s22.log
run() { echo "== $*"; env "$@" timeout 200 python tools/bench_kernels.py --shape $SHAPE --prec ${PREC:-f64} --reps 3 --only3d --fmt ${FMT:-Z} 2>&1 | tail -7; }
SHAPE=510,510,510 FMT=ZX run D2D_ANY_BIG=1
SHAPE=360,360,360 run D2D_ANY_BIG=1
SHAPE=544,416,352 run D2D_ANY_BIG=1
SHAPE=1000,1000,1000 run D2D_ANY_BIG=1
SHAPE=513,513,513 run D2D_ANY_BIG=1
SHAPE=1025,1025,257 run D2D_ANY_BIG=1
PREC=f32 SHAPE=510,510,510 run D2D_ANY_BIG=1
