#!/bin/bash
set -x
mkdir -p gpurun_out; O=gpurun_out
date +%T
timeout 300 python -u -m pytest tests/test_gpu_fft1d.py tests/test_gpu_multiple_grids.py -m gpu -q --maxfail=10 -p no:cacheprovider --timeout=300 > $O/pytest_s10.log 2>&1; tail -5 $O/pytest_s10.log
date +%T
L=$PWD/2decomp-fft_b200/lib
bk() { name=$1; shift; echo "== $name"; timeout 200 env "$@" python -u tools/bench_kernels.py --reps 5 --only3d $KARGS > $O/k_$name.txt 2>&1; cat $O/k_$name.txt | grep -v "^$"; }
bk st_default X=1
bk st_cs D2D_B200_LIB=$L/libd2dfft_b200_cs.so
bk st_cg D2D_B200_LIB=$L/libd2dfft_b200_cg.so
KARGS="--n 2048 --prec f32 --fmt X --reps 2" bk f32_2048_new X=1
KARGS="--n 2048 --prec f32 --fmt X --reps 2" bk f32_2048_old D2D_B200_LIB=$L/libd2dfft_b200_old2048.so
KARGS="--shape 2048,512,512 --prec f32 --fmt ZX --reps 3" bk f32_2048s_new X=1
KARGS="--shape 2048,512,512 --prec f32 --fmt ZX --reps 3" bk f32_2048s_old D2D_B200_LIB=$L/libd2dfft_b200_old2048.so
date +%T
