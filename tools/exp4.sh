echo "== tests with wide tiles"; D2D_V2_ROWBYTES=128 D2D_V2_ROWBYTES_TILEOUT=128 timeout 900 python -m pytest tests/test_gpu_fft1d.py tests/test_gpu_fft3d.py -m gpu -x -q 2>&1 | tail -4
echo "== 64/64"; python tools/bench_kernels.py --n 1024 --reps 5 2>&1 | grep -E "Z:"
echo "== 64/128 (tile-out stages wide)"; D2D_V2_ROWBYTES_TILEOUT=128 python tools/bench_kernels.py --n 1024 --reps 5 2>&1 | grep -E "Z:|X:"
echo "== 128/128"; D2D_V2_ROWBYTES=128 D2D_V2_ROWBYTES_TILEOUT=128 python tools/bench_kernels.py --n 1024 --reps 5 2>&1 | grep -E "Z:|X:"
echo "== 128/128 skip3 (compute only)"; D2D_DEBUG_SKIP=3 D2D_V2_ROWBYTES=128 D2D_V2_ROWBYTES_TILEOUT=128 python tools/bench_kernels.py --n 1024 --reps 5 2>&1 | grep -E "Z:"
