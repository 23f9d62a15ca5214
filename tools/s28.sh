#!/bin/bash
mkdir -p gpurun_out
timeout 60 python -c "
import __graft_entry__ as g, torch
pkg, orc = g.package(), g.oracle()
torch.cuda.set_device(0)
g._smoke_case(pkg, orc, (96, 80, 48), False)
g._smoke_case(pkg, orc, (34, 26, 22), False)
" > gpurun_out/smoke_s28.log 2>&1; tail -4 gpurun_out/smoke_s28.log
