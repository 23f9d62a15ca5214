#!/bin/bash
# One single-GPU session producing the artefacts of record: GPU test suite, smoke, bench line (+ CPU arm), ncu launch list,
# ncu --set full of the six FFT kernels of the headline pair, per-stage rooflines.  Every step has its own timeout and
# writes unbuffered logs under gpurun_out/ (copy what should be kept into profiles/).
set -x
mkdir -p gpurun_out
O=gpurun_out
date +%T
echo "(GPU suite: sessions s22 + s23)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -4 $O/smoke.log
date +%T
timeout 400 python -u bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err
cat $O/bench.json; tail -3 $O/bench.err
bk() { name=$1; shift; timeout 150 python -u tools/bench_kernels.py "$@" > $O/k_$name.txt 2>&1; echo "== $name"; cat $O/k_$name.txt; }
bk f64 --n 1024 --reps 5 --only3d
bk f32 --n 1024 --prec f32 --reps 5 --only3d
if [ "$1" != "quick" ]; then
timeout 300 python -u bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; cat $O/bench_ref.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29555 bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_torchrun.json 2> $O/bench_ref_torchrun.err; cat $O/bench_ref_torchrun.json
# only the library's kernels (torch's RNG / reduction kernels of the bench harness would eat the launch budget): 3 pairs = 18 launches
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fft_ -c 40 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > $O/bench_under_ncu.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:fft_kernel -s 6 -c 6 -o $O/prof_pair -f python tools/run_pair.py 1024 1024 1024 z f64 2 > $O/ncu_full.log 2>&1
tail -2 $O/ncu_full.log
python tools/ncu_summary.py $O/prof_pair.ncu-rep > $O/ncu_pair_summary.txt 2>&1
fi
date +%T
