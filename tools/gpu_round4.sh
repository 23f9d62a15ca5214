#!/bin/bash
# Single-GPU session: tests on the new default (one-exchange plans), bench line, row-width A/B runs.
set -x
mkdir -p gpurun_out
O=gpurun_out
R16=$PWD/2decomp-fft_b200/lib/libd2dfft_b200_r16.so
date +%T
timeout 300 python -u -m pytest tests/test_cabi.py tests/test_golden.py tests/test_gpu_fft1d.py tests/test_gpu_fft3d.py -m gpu -q --maxfail=10 -p no:cacheprovider --timeout=90 > $O/pytest_gpu_core.log 2>&1; tail -4 $O/pytest_gpu_core.log
timeout 300 python -u -m pytest tests/test_gpu_fft_any.py -m gpu -q --maxfail=10 -p no:cacheprovider --timeout=90 > $O/pytest_gpu_any.log 2>&1; tail -4 $O/pytest_gpu_any.log
timeout 400 python -u -m pytest tests/test_gpu_configs.py -m gpu -v --maxfail=10 -p no:cacheprovider --timeout=200 > $O/pytest_gpu_configs.log 2>&1; tail -12 $O/pytest_gpu_configs.log
date +%T
timeout 300 python -u bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err
cat $O/bench.json; tail -3 $O/bench.err
date +%T
bk() { name=$1; shift; timeout 150 python -u tools/bench_kernels.py "$@" > $O/k_$name.txt 2>&1; echo "== $name"; cat $O/k_$name.txt; }
bk f64_default --n 1024 --reps 5 --only3d
D2D_V2_ROWBYTES=128 bk f64_all128 --n 1024 --reps 5 --only3d
D2D_V2_ROWBYTES_R2C=128 bk f64_r2c128 --n 1024 --reps 5 --only3d
D2D_V2_ROWBYTES_TILEOUT=64 bk f64_all64 --n 1024 --reps 5 --only3d --fmt Z
bk f32_default --n 1024 --prec f32 --reps 5 --only3d
D2D_V2_ROWBYTES_TILEOUT=128 bk f32_tout128 --n 1024 --prec f32 --reps 5 --only3d
D2D_V2_ROWBYTES=128 D2D_V2_ROWBYTES_TILEOUT=128 bk f32_all128 --n 1024 --prec f32 --reps 5 --only3d
D2D_B200_LIB=$R16 bk f32_r16 --n 1024 --prec f32 --reps 5 --only3d --fmt Z
bk f64_512 --n 512 --reps 10 --only3d
D2D_V2_ROWBYTES=128 bk f64_512_all128 --n 512 --reps 10 --only3d
bk f32_2048 --n 2048 --prec f32 --reps 3 --only3d --fmt X
date +%T
ls -la $O
