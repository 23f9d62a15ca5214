timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:fft_kernel -s 6 -c 6 -o gpurun_out/prof_pair_v2 -f python tools/run_pair.py 1024 1024 1024 z f64 2 > gpurun_out/ncu_full_v2.log 2>&1
tail -2 gpurun_out/ncu_full_v2.log
for m in 1 2 3; do echo "== D2D_DEBUG_SKIP=$m"; D2D_DEBUG_SKIP=$m python tools/bench_kernels.py --n 1024 --reps 3 2>&1 | grep -E "Z:"; done
