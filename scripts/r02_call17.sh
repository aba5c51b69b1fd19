#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r02_wide_probe2.txt
: > $O
for pair in 0 1; do SPGNN_WIDE_PAIR=$pair timeout -k 5 200 python scripts/wide_pair_probe.py >> $O 2>&1; echo "rc=$?" >> $O; done
cat $O
echo "== nt pair (one arrival per warp)" > gpurun_out/r02_nt_pair2.txt
timeout -k 5 150 python scripts/planes_check.py --bench >> gpurun_out/r02_nt_pair2.txt 2>&1; echo "rc=$?" >> gpurun_out/r02_nt_pair2.txt
echo "== single" >> gpurun_out/r02_nt_pair2.txt
SPGNN_NT_PAIR=0 timeout -k 5 150 python scripts/planes_check.py --bench >> gpurun_out/r02_nt_pair2.txt 2>&1
grep "worst\|gat\|pgnn\|head\|==\|rc=" gpurun_out/r02_nt_pair2.txt
