#!/bin/bash
# warp-convergent UMMA issue loops: correctness + per-layer timing of every GEMM kernel, then the step
mkdir -p gpurun_out
O=gpurun_out/r02_umma_issue.txt
echo "== planes_check (pair default)" > $O
timeout -k 5 200 python scripts/planes_check.py --bench >> $O 2>&1; echo "rc=$?" >> $O
echo "== planes_check (SPGNN_NT_PAIR=0)" >> $O
SPGNN_NT_PAIR=0 timeout -k 5 200 python scripts/planes_check.py --bench >> $O 2>&1; echo "rc=$?" >> $O
echo "== wide probe" >> $O
for pair in 0 1; do SPGNN_WIDE_PAIR=$pair timeout -k 5 200 python scripts/wide_pair_probe.py >> $O 2>&1; echo "rc=$?" >> $O; done
timeout -k 5 200 python scripts/wide_check.py >> $O 2>&1; echo "rc=$?" >> $O
grep -v "^M=" $O
for rep in 1 2; do
  timeout -k 5 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-small --stream-steps 0 > gpurun_out/issue.json 2>gpurun_out/issue.err
  python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/issue.json').read().strip().splitlines()[-1])
    print('step %.2f ms'%d['ms_per_step'], 'infer %.2f ms'%d['infer']['ms_per_step'], 'loss', d['config']['loss'], 'sm_mhz', d['clocks'].get('sm_mhz'), d['kernel_time_shares'])
except Exception as e:
    print('failed', e, open('gpurun_out/issue.err').read()[-800:])
PY
done
