for m in 0 1 2 3; do echo "== D2D_DEBUG_SKIP=$m"; D2D_DEBUG_SKIP=$m python tools/bench_kernels.py --n 1024 --reps 3 2>&1 | grep -E "Z:|X:"; done
echo "== wide real tiles"; D2D_WIDE_REAL_TILES=1 python tools/bench_kernels.py --n 1024 --reps 3 2>&1 | grep -E "Z:|X:"
