#!/bin/bash
# CTA-pair (cta_group::2) NT GEMM: correctness against fp64 and per-layer timing vs the single-CTA kernel
mkdir -p gpurun_out
echo "== pair" > gpurun_out/r02_nt_pair.txt
SPGNN_NT_PAIR=1 timeout -k 5 150 python scripts/planes_check.py --bench >> gpurun_out/r02_nt_pair.txt 2>&1; echo "rc=$?" >> gpurun_out/r02_nt_pair.txt
nvidia-smi --query-gpu=name,memory.used --format=csv >> gpurun_out/r02_nt_pair.txt
echo "== single" >> gpurun_out/r02_nt_pair.txt
timeout -k 5 300 python scripts/planes_check.py --bench >> gpurun_out/r02_nt_pair.txt 2>&1; echo "rc=$?" >> gpurun_out/r02_nt_pair.txt
cat gpurun_out/r02_nt_pair.txt
