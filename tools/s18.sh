#!/bin/bash
set -x
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -u -m pytest tests/test_gpu_fft_any.py tests/test_gpu_multiple_grids.py -m gpu -q --maxfail=20 -p no:cacheprovider --timeout=300 > $O/pytest_s18.log 2>&1; tail -15 $O/pytest_s18.log
for big in 0 1; do
for s in 510,510,510 544,416,352 1000,1000,1000 360,360,360; do
  D2D_ANY_BIG=$big timeout 200 python tools/bench_kernels.py --shape $s --prec f64 --reps 3 --only3d --fmt Z > $O/k_any_${s}_big$big.txt 2>&1; echo "== $s big=$big"; tail -9 $O/k_any_${s}_big$big.txt
done
done
D2D_ANY_BIG=1 timeout 200 python tools/bench_kernels.py --shape 510,510,510 --prec f32 --reps 3 --only3d --fmt Z 2>&1 | tail -9
D2D_ANY_BIG=0 timeout 200 python tools/bench_kernels.py --shape 510,510,510 --prec f32 --reps 3 --only3d --fmt Z 2>&1 | tail -9
