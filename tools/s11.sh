#!/bin/bash
set -x
mkdir -p gpurun_out; O=gpurun_out
date +%T
timeout 900 python -u -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider --timeout=300 > $O/pytest_s11.log 2>&1; tail -12 $O/pytest_s11.log
date +%T
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_s11.log 2>&1; tail -4 $O/smoke_s11.log
date +%T
