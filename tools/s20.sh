#!/bin/bash
set -x
mkdir -p gpurun_out; O=gpurun_out
# any-length kernel under ncu: 510^3 fp64 pair (small and big build), launches 7.. (after one warm pair)
for big in 0 1; do
D2D_ANY_BIG=$big timeout 400 ncu --set full --clock-control none --import-source on -k regex:fft_any -s 6 -c 3 -o $O/prof_any510_big$big -f python tools/run_pair.py 510 510 510 z f64 2 > $O/ncu_any_$big.log 2>&1
tail -2 $O/ncu_any_$big.log
python tools/ncu_summary.py $O/prof_any510_big$big.ncu-rep > $O/ncu_any510_big${big}_summary.txt 2>&1
for k in 1 2 3; do python tools/ncu_stalls.py $O/prof_any510_big$big.ncu-rep $k 30 > $O/ncu_any510_big${big}_stalls_$k.txt 2>&1; done
done
ls -la $O/*.ncu-rep
