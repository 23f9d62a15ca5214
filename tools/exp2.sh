timeout 900 python -m pytest tests/test_gpu_fft1d.py -m gpu -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/bench_kernels.py --n 1024 --reps 5 2>&1 | tail -16
for p in 0 2; do echo "== L2PROMO=$p"; D2D_TMA_L2PROMO=$p timeout 300 python tools/bench_kernels.py --n 1024 --reps 5 2>&1 | grep "Z:"; done
