#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r02_tn_split_sources.txt
: > $O
for sp in 0 1; do
  echo "== SPGNN_TN_SPLIT_SOURCES=$sp" >> $O
  SPGNN_TN_SPLIT_SOURCES=$sp timeout -k 5 200 python scripts/planes_check.py --bench 2>&1 | grep "worst\|gat0\|gat1\|M=40100\|M=38001\|M=1000 " >> $O
done
for rep in 1 2; do for sp in 0 1; do
  SPGNN_TN_SPLIT_SOURCES=$sp timeout -k 5 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-small --stream-steps 0 > gpurun_out/sp.json 2>gpurun_out/sp.err
  python - $sp <<'PY' >> gpurun_out/r02_tn_split_sources.txt
import json, sys
try:
    d=json.loads(open('gpurun_out/sp.json').read().strip().splitlines()[-1])
    print('split sources', sys.argv[1], 'step %.2f ms'%d['ms_per_step'], 'loss', d['config']['loss'], 'roofline', round(d['roofline']['frac'],4), round(d['roofline']['avg_ms'],3), d['kernel_time_shares'])
except Exception as e:
    print('failed', e, open('gpurun_out/sp.err').read()[-600:])
PY
done; done
cat $O
timeout -k 5 600 python -m pytest tests -m gpu -q --timeout 600 -x -k "fp64 or fullsize or gradients or train_mode_step" 2>&1 | tail -3
