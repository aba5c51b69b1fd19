#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r02_wide_probe.txt
: > $O
for pair in 0 1; do SPGNN_WIDE_PAIR=$pair timeout -k 5 200 python scripts/wide_pair_probe.py >> $O 2>&1; echo "rc=$?" >> $O; done
cat $O
timeout -k 5 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "dx_gemm" 2>&1 | tail -3
