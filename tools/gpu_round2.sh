#!/bin/bash
# One single-GPU session: GPU tests, bench line (+ reference arm), ncu launch list, ncu --set full of the six FFT kernels of
# the headline pair, per-kernel rooflines for both formats / precisions, any-length kernel sample, one experiment.
set -x
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > $O/pytest_gpu.log 2>&1
tail -15 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err
cat $O/bench.json; tail -3 $O/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
cat $O/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > $O/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fft_kernel -s 6 -c 6 -o $O/prof_pair -f python tools/run_pair.py 1024 1024 1024 z f64 2 > $O/ncu_full.log 2>&1
tail -2 $O/ncu_full.log
python tools/bench_kernels.py --n 1024 --reps 5 > $O/kernels_1024_f64.txt 2>&1; cat $O/kernels_1024_f64.txt
D2D_V2_ROWBYTES_R2C=128 python tools/bench_kernels.py --n 1024 --reps 5 --only3d > $O/kernels_1024_f64_r2c128.txt 2>&1; cat $O/kernels_1024_f64_r2c128.txt
python tools/bench_kernels.py --n 1024 --prec f32 --reps 5 --only3d > $O/kernels_1024_f32.txt 2>&1; cat $O/kernels_1024_f32.txt
python tools/bench_kernels.py --n 2048 --prec f32 --reps 3 --only3d --fmt X > $O/kernels_2048_f32.txt 2>&1; cat $O/kernels_2048_f32.txt
python tools/bench_kernels.py --shape 510,510,510 --reps 3 --only3d > $O/kernels_510_any.txt 2>&1; cat $O/kernels_510_any.txt
python tools/bench_kernels.py --shape 680,520,440 --reps 3 --only3d --fmt X > $O/kernels_680_any.txt 2>&1; cat $O/kernels_680_any.txt
ls -la $O
