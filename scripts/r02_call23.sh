#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/r02_pytest_23.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_23.log
O=gpurun_out/r02_aggx_lanes.txt
: > $O
for rep in 1 2; do for lanes in 32 16; do
  SPGNN_AGGX_LANES=$lanes timeout -k 5 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-small --stream-steps 0 > gpurun_out/ax.json 2>gpurun_out/ax.err
  python - $lanes <<'PY' >> gpurun_out/r02_aggx_lanes.txt
import json, sys
try:
    d=json.loads(open('gpurun_out/ax.json').read().strip().splitlines()[-1])
    print('aggx lanes per node', sys.argv[1], 'step %.2f ms'%d['ms_per_step'], 'infer %.2f'%d['infer']['ms_per_step'], 'loss', d['config']['loss'], d['kernel_time_shares'])
except Exception as e:
    print('failed', e, open('gpurun_out/ax.err').read()[-600:])
PY
done; done
cat $O
