#!/bin/bash
set -x
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -u -m pytest tests/test_gpu_fft_any.py -m gpu -q --maxfail=20 -p no:cacheprovider --timeout=300 > $O/pytest_s19.log 2>&1; tail -5 $O/pytest_s19.log
for big in 0 1; do
for s in 510,510,510 544,416,352 360,360,360; do
  echo "== $s big=$big"; D2D_ANY_BIG=$big timeout 200 python tools/bench_kernels.py --shape $s --prec f64 --reps 3 --only3d --fmt Z 2>&1 | tail -7
done
done
echo "== 1000 4 lines"; timeout 200 python tools/bench_kernels.py --shape 1000,1000,1000 --prec f64 --reps 2 --only3d --fmt Z 2>&1 | tail -7
echo "== 1000 2 lines"; D2D_ANY_MIN_ROW_BYTES=32 timeout 200 python tools/bench_kernels.py --shape 1000,1000,1000 --prec f64 --reps 2 --only3d --fmt Z 2>&1 | tail -7
echo "== 510 f32"; timeout 200 python tools/bench_kernels.py --shape 510,510,510 --prec f32 --reps 3 --only3d --fmt Z 2>&1 | tail -7
