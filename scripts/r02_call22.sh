#!/bin/bash
# L2 prefetch of the GEMM operands a few k-blocks ahead of their TMA loads: per-layer timing and the step
mkdir -p gpurun_out
O=gpurun_out/r02_gemm_prefetch.txt
: > $O
for pf in 0 4 8 16; do
  echo "== SPGNN_GEMM_PREFETCH=$pf" >> $O
  SPGNN_GEMM_PREFETCH=$pf timeout -k 5 200 python scripts/planes_check.py --bench 2>&1 | grep "worst\|gat\|pgnn\|head" >> $O
  SPGNN_GEMM_PREFETCH=$pf timeout -k 5 200 python scripts/wide_check.py --big 2>&1 | grep "worst\|wide:" >> $O
done
for rep in 1 2; do for pf in 0 8; do
  SPGNN_GEMM_PREFETCH=$pf timeout -k 5 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-small --stream-steps 0 > gpurun_out/pf.json 2>gpurun_out/pf.err
  python - $pf <<'PY' >> gpurun_out/r02_gemm_prefetch.txt
import json, sys
try:
    d=json.loads(open('gpurun_out/pf.json').read().strip().splitlines()[-1])
    print('prefetch', sys.argv[1], 'step %.2f ms'%d['ms_per_step'], 'infer %.2f'%d['infer']['ms_per_step'], 'loss', d['config']['loss'], 'sm_mhz', d['clocks'].get('sm_mhz'), d['kernel_time_shares'])
except Exception as e:
    print('failed', e, open('gpurun_out/pf.err').read()[-600:])
PY
done; done
cat $O
