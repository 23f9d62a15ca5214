#!/bin/bash
# Single-GPU session, most important artefacts first, every step under its own timeout, logs unbuffered.
set -x
mkdir -p gpurun_out
O=gpurun_out
SAFE=$PWD/2decomp-fft_b200/lib/libd2dfft_b200_safe.so
R32=$PWD/2decomp-fft_b200/lib/libd2dfft_b200_r32.so
date +%T
timeout 300 python -u bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err
cat $O/bench.json; tail -3 $O/bench.err
if ! grep -q '"metric"' $O/bench.json; then
  D2D_B200_LIB=$SAFE timeout 300 python -u bench.py --steps 10 --warmup 3 > $O/bench_safe.json 2> $O/bench_safe.err; cat $O/bench_safe.json; tail -3 $O/bench_safe.err
fi
date +%T
timeout 120 python -u -m pytest tests/test_gpu_fft1d.py -q -x -p no:cacheprovider -k "512 or 1024 or 2048" --timeout=60 > $O/pytest_quick.log 2>&1; tail -3 $O/pytest_quick.log
timeout 200 python -u bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; cat $O/bench_ref.json
date +%T
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > $O/bench_under_ncu.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:fft_kernel -s 6 -c 6 -o $O/prof_pair -f python tools/run_pair.py 1024 1024 1024 z f64 2 > $O/ncu_full.log 2>&1
tail -2 $O/ncu_full.log
date +%T
timeout 120 python -u tools/bench_kernels.py --n 1024 --reps 5 > $O/kernels_1024_f64.txt 2>&1; cat $O/kernels_1024_f64.txt
D2D_B200_LIB=$SAFE timeout 120 python -u tools/bench_kernels.py --n 1024 --reps 5 --only3d > $O/kernels_1024_f64_safe.txt 2>&1; cat $O/kernels_1024_f64_safe.txt
D2D_B200_LIB=$R32 timeout 120 python -u tools/bench_kernels.py --n 1024 --reps 5 --only3d > $O/kernels_1024_f64_r32.txt 2>&1; cat $O/kernels_1024_f64_r32.txt
D2D_B200_LIB=$R32 timeout 120 python -u -m pytest tests/test_gpu_fft1d.py -q -x -p no:cacheprovider -k "512 or 1024" --timeout=60 > $O/pytest_r32.log 2>&1; tail -3 $O/pytest_r32.log
for a in "64 1 4 8388608" "128 1 4 8388608" "64 2 4 8388608" "128 2 4 8388608" "64 1 4 8192" "128 2 4 8192"; do timeout 60 tools/micro/membench $a; done > $O/membench_pitch.txt 2>&1; cat $O/membench_pitch.txt
date +%T
timeout 300 python -u -m pytest tests/test_cabi.py tests/test_golden.py tests/test_gpu_fft1d.py tests/test_gpu_fft3d.py -m gpu -q --maxfail=10 -p no:cacheprovider --timeout=90 > $O/pytest_gpu_core.log 2>&1; tail -6 $O/pytest_gpu_core.log
date +%T
timeout 300 python -u -m pytest tests/test_gpu_fft_any.py -m gpu -q --maxfail=10 -p no:cacheprovider --timeout=90 > $O/pytest_gpu_any.log 2>&1; tail -6 $O/pytest_gpu_any.log
date +%T
timeout 400 python -u -m pytest tests/test_gpu_configs.py -m gpu -v --maxfail=10 -p no:cacheprovider --timeout=150 > $O/pytest_gpu_configs.log 2>&1; tail -12 $O/pytest_gpu_configs.log
date +%T
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 120 python -u tools/bench_kernels.py --n 1024 --prec f32 --reps 5 --only3d > $O/kernels_1024_f32.txt 2>&1; cat $O/kernels_1024_f32.txt
timeout 120 python -u tools/bench_kernels.py --shape 510,510,510 --reps 3 --only3d > $O/kernels_510_any.txt 2>&1; cat $O/kernels_510_any.txt
timeout 150 python -u tools/bench_kernels.py --n 2048 --prec f32 --reps 3 --only3d --fmt X > $O/kernels_2048_f32.txt 2>&1; cat $O/kernels_2048_f32.txt
date +%T
ls -la $O
