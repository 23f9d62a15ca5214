#!/bin/bash
# 1-GPU session: swizzled merged landing A/B, e2e spans
set -x
mkdir -p gpurun_out; O=gpurun_out
date +%T
timeout 300 python -u -m pytest tests/test_gpu_fft1d.py tests/test_gpu_fft3d.py tests/test_gpu_configs.py -m gpu -q --maxfail=10 -p no:cacheprovider --timeout=300 > $O/pytest_gpu_swz.log 2>&1; tail -5 $O/pytest_gpu_swz.log
date +%T
bk() { name=$1; shift; echo "== $name"; timeout 150 env "$@" python -u tools/bench_kernels.py --reps 5 --only3d > $O/k_$name.txt 2>&1; cat $O/k_$name.txt | grep -v "^$"; }
bk swz1 D2D_V2_SWIZZLE=1 X=1
bk swz0 D2D_V2_SWIZZLE=0
bk swz1_m2 D2D_V2_SWIZZLE=1 D2D_V2_MERGE=2
bk swz0_m2 D2D_V2_SWIZZLE=0 D2D_V2_MERGE=2
echo "== f32"
timeout 150 env D2D_V2_SWIZZLE=1 python -u tools/bench_kernels.py --reps 5 --only3d --prec f32 > $O/k_f32_swz1.txt 2>&1; cat $O/k_f32_swz1.txt
timeout 150 env D2D_V2_SWIZZLE=0 python -u tools/bench_kernels.py --reps 5 --only3d --prec f32 > $O/k_f32_swz0.txt 2>&1; cat $O/k_f32_swz0.txt
timeout 150 env D2D_V2_SWIZZLE=1 D2D_V2_MERGE=2 python -u tools/bench_kernels.py --reps 5 --only3d --prec f32 > $O/k_f32_swz1_m2.txt 2>&1; cat $O/k_f32_swz1_m2.txt
date +%T
timeout 300 python -u bench.py --steps 5 --warmup 3 --no-cpu > $O/bench_s3.json 2> $O/bench_s3.err; python -c "
import json; j=json.loads(open('$O/bench_s3.json').read().strip().splitlines()[-1]); print(j['ms_per_step'], j['e2e'])"; tail -3 $O/bench_s3.err
date +%T
