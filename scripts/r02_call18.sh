#!/bin/bash
# source-level ncu of the wide (output layer) GEMM at the bench size: where do its warps wait?
mkdir -p gpurun_out
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:'wide_kernel' -o /tmp/wide -f python scripts/wide_check.py --once > gpurun_out/r02_wide_ncu.log 2>&1
ls -la /tmp/wide.ncu-rep
python scripts/ncu_stalls.py /tmp/wide.ncu-rep 0 45 > gpurun_out/r02_wide_stalls_mode0.txt 2>&1
python scripts/ncu_stalls.py /tmp/wide.ncu-rep 1 45 > gpurun_out/r02_wide_stalls_mode1.txt 2>&1
python scripts/ncu_summary.py /tmp/wide.ncu-rep 4 > gpurun_out/r02_wide_ncu_summary.txt 2>&1
head -50 gpurun_out/r02_wide_stalls_mode0.txt
