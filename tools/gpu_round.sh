#!/bin/bash
# One GPU session: bench line, ncu launch list, one ncu --set full capture of the FFT kernels, GPU tests.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > gpurun_out/clocks.csv &
SMI=$!
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json
tail -5 gpurun_out/bench.err
kill $SMI
python tools/bench_kernels.py --n 1024 --reps 5 > gpurun_out/kernels_1024.txt 2>&1
cat gpurun_out/kernels_1024.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fft_kernel -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:fft_kernel -s 6 -c 6 -o gpurun_out/prof_pair -f python tools/run_pair.py 1024 1024 1024 z f64 2 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cat gpurun_out/bench_ref.json
