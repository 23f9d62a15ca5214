#!/bin/bash
# 1-GPU session: new tests, lock-step clusters A/B, paired stores, ncu of the X-format pair
set -x
mkdir -p gpurun_out; O=gpurun_out
date +%T
timeout 600 python -u -m pytest tests/test_gpu_multiple_grids.py tests/test_c_example.py tests/test_gpu_configs.py tests/test_gpu_fft1d.py tests/test_gpu_fft3d.py -m gpu -q --maxfail=10 -p no:cacheprovider --timeout=300 > $O/pytest_s9.log 2>&1; tail -15 $O/pytest_s9.log
date +%T
bk() { name=$1; shift; echo "== $name"; timeout 150 env "$@" python -u tools/bench_kernels.py --reps 5 --only3d > $O/k_$name.txt 2>&1; cat $O/k_$name.txt | grep -v "^$"; }
bk cl1 D2D_CLUSTER=1
bk cl2 D2D_CLUSTER=2
bk cl4 D2D_CLUSTER=4
echo "== f32 cl1/cl2"
timeout 150 env D2D_CLUSTER=1 python -u tools/bench_kernels.py --reps 5 --only3d --prec f32 > $O/k_f32_cl1.txt 2>&1; cat $O/k_f32_cl1.txt
timeout 150 env D2D_CLUSTER=2 python -u tools/bench_kernels.py --reps 5 --only3d --prec f32 > $O/k_f32_cl2.txt 2>&1; cat $O/k_f32_cl2.txt
date +%T
timeout 500 ncu --set full --clock-control none --import-source on -k regex:fft_kernel -s 6 -c 6 -o $O/prof_pair_x -f python tools/run_pair.py 1024 1024 1024 x f64 2 > $O/ncu_full_x.log 2>&1
tail -2 $O/ncu_full_x.log
python tools/ncu_summary.py $O/prof_pair_x.ncu-rep > $O/ncu_pair_x_summary.txt 2>&1; head -5 $O/ncu_pair_x_summary.txt
date +%T
