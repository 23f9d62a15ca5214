#!/bin/bash
set -x
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -u -m pytest tests/test_gpu_fft_mixed.py -m gpu -q --maxfail=20 -p no:cacheprovider --timeout=300 > $O/pytest_s17.log 2>&1; tail -30 $O/pytest_s17.log
timeout 600 python -u -m pytest tests/test_gpu_fft1d.py -m gpu -q --maxfail=20 -p no:cacheprovider --timeout=300 -k 'chirp or large_prime or unsupported or long_lines' > $O/pytest_s17b.log 2>&1; tail -30 $O/pytest_s17b.log
for s in 768 384 640; do timeout 200 python tools/bench_kernels.py --n $s --prec f64 --reps 3 > $O/k_mixed_${s}_f64.txt 2>&1; tail -25 $O/k_mixed_${s}_f64.txt; done
timeout 200 python tools/bench_kernels.py --n 768 --prec f32 --reps 3 > $O/k_mixed_768_f32.txt 2>&1; tail -25 $O/k_mixed_768_f32.txt
timeout 200 python tools/bench_kernels.py --n 510 --prec f64 --reps 3 --only3d > $O/k_any_510_f64.txt 2>&1; tail -8 $O/k_any_510_f64.txt
