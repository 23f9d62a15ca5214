timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/bench_kernels.py --n 1024 --reps 5 2>&1 | grep -E "Z:|X:"
echo "== skip3"; D2D_DEBUG_SKIP=3 python tools/bench_kernels.py --n 1024 --reps 5 2>&1 | grep -E "Z:"
echo "== 512"; python tools/bench_kernels.py --n 512 --reps 10 2>&1 | grep -E "Z:|X:"
echo "== f32 1024"; python tools/bench_kernels.py --n 1024 --prec f32 --reps 5 2>&1 | grep -E "Z:|X:"
