set -x
O=gpurun_out; mkdir -p $O
timeout 400 python -u -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider --timeout=200 > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
bk() { name=$1; shift; timeout 100 python -u tools/bench_kernels.py "$@" > $O/k_$name.txt 2>&1; echo "== $name"; cat $O/k_$name.txt; }
bk any768_128 --shape 768,768,768 --reps 3 --only3d --fmt Z
D2D_ANY_ROW_BYTES=64 bk any768_64 --shape 768,768,768 --reps 3 --only3d --fmt Z
D2D_ANY_ROW_BYTES=64 bk any510_64 --shape 510,510,510 --reps 3 --only3d --fmt Z
timeout 100 python -u tools/bench_transpose.py 512 2 2 > $O/transpose_512.txt 2>&1; cat $O/transpose_512.txt
