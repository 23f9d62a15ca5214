#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
B=tools/micro/nvlbench
{
$B 0 0 0
for run in 64 128 256 512 2048; do $B 1 $run 296 256; done
for run in 128 256 512; do $B 1 $run 592 256; done
for run in 128 256 512 2048; do $B 2 $run 296 256; done
$B 2 512 1184 256
for p in 4096 16384 32768; do for bl in 8 16 32 148; do $B 3 $p $bl 32 6; done; done
$B 3 16384 16 32 12
$B 3 8192 32 32 12
$B 3 32768 16 32 4
} 2>&1 | tee $O/nvlbench.txt
