#!/bin/bash
set -x
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -u -m pytest tests/test_gpu_multiproc.py tests/test_gpu_fft3d.py -m gpu -q --maxfail=20 -p no:cacheprovider --timeout=600 -k "multiproc or cap or shared or process_grid" > $O/pytest_s23.log 2>&1; tail -5 $O/pytest_s23.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_s23.log 2>&1; tail -6 $O/smoke_s23.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fft_any -s 6 -c 3 -o $O/prof_any510_async -f python tools/run_pair.py 510 510 510 z f64 2 > $O/ncu_any_async.log 2>&1
python tools/ncu_summary.py $O/prof_any510_async.ncu-rep > $O/ncu_any510_async_summary.txt 2>&1
for k in 1 2 3; do python tools/ncu_stalls.py $O/prof_any510_async.ncu-rep $k 30 > $O/ncu_any510_async_stalls_$k.txt 2>&1; done
head -30 $O/ncu_any510_async_stalls_2.txt
