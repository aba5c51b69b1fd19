#!/bin/bash
# A/B of two library builds on ONE box: spgnn_b200/lib_base.so (baseline) vs the tree's libspgnn_b200.so
mkdir -p gpurun_out
for rep in 1 2; do
for which in base new; do
  if [ $which = base ]; then export SPGNN_B200_LIB=$PWD/spgnn_b200/lib_base.so; else unset SPGNN_B200_LIB; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --stream-steps 0 > gpurun_out/ab_$which.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open('gpurun_out/ab_$which.json').read().strip().splitlines()[-1])
ra=d['roofline_agg']
print('$which', 'step %.2f ms'%d['ms_per_step'], 'agg fwd %.3f ms (%.3f)'%(ra['fwd']['avg_ms'],ra['fwd']['frac']), 'bwd %.3f ms (%.3f)'%(ra['bwd']['avg_ms'],ra['bwd']['frac']), d['kernel_time_shares'])
PY
done; done
