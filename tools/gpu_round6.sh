#!/bin/bash
# Single-GPU session: merged-landing r2c A/B, any-length kernel v2, fp32 defaults, artefacts of record.
set -x
mkdir -p gpurun_out
O=gpurun_out
date +%T
timeout 300 python -u -m pytest tests/test_cabi.py tests/test_golden.py tests/test_gpu_fft1d.py tests/test_gpu_fft3d.py -m gpu -q --maxfail=10 -p no:cacheprovider --timeout=90 > $O/pytest_gpu_core.log 2>&1; tail -4 $O/pytest_gpu_core.log
timeout 300 python -u -m pytest tests/test_gpu_fft_any.py -m gpu -q --maxfail=10 -p no:cacheprovider --timeout=90 > $O/pytest_gpu_any.log 2>&1; tail -4 $O/pytest_gpu_any.log
timeout 400 python -u -m pytest tests/test_gpu_configs.py -m gpu -q --maxfail=10 -p no:cacheprovider --timeout=200 > $O/pytest_gpu_configs.log 2>&1; tail -3 $O/pytest_gpu_configs.log
D2D_V2_MERGE=2 timeout 200 python -u -m pytest tests/test_gpu_fft1d.py -m gpu -q -k "1024" --maxfail=10 -p no:cacheprovider --timeout=90 > $O/pytest_gpu_merge2.log 2>&1; tail -3 $O/pytest_gpu_merge2.log
date +%T
bk() { name=$1; shift; timeout 150 python -u tools/bench_kernels.py "$@" > $O/k_$name.txt 2>&1; echo "== $name"; cat $O/k_$name.txt; }
D2D_V2_MERGE=0 bk f64_merge0 --n 1024 --reps 5 --only3d --fmt Z
D2D_V2_MERGE=1 bk f64_merge1 --n 1024 --reps 5 --only3d
D2D_V2_MERGE=2 bk f64_merge2 --n 1024 --reps 5 --only3d --fmt Z
bk f32_default --n 1024 --prec f32 --reps 5 --only3d
bk any_510 --shape 510,510,510 --reps 3 --only3d --fmt Z
bk any_768 --shape 768,768,768 --reps 3 --only3d --fmt Z
bk any_17x --shape 544,416,352 --reps 3 --only3d --fmt X
date +%T
timeout 300 python -u bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err
cat $O/bench.json; tail -3 $O/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > $O/bench_under_ncu.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:fft_kernel -s 6 -c 6 -o $O/prof_pair -f python tools/run_pair.py 1024 1024 1024 z f64 2 > $O/ncu_full.log 2>&1
tail -2 $O/ncu_full.log
date +%T
