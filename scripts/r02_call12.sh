#!/bin/bash
# parity suite with the new tree-backward launch shape + mask folded into the dX GEMM, then A/B of the fold
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/r02_pytest_12.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest_12.log
O=gpurun_out/r02_fold_mask_ab.txt
: > $O
for rep in 1 2; do
for fold in 0 1; do
  SPGNN_FOLD_MASK=$fold timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-small --stream-steps 0 > gpurun_out/fm_$fold.json 2>gpurun_out/fm_$fold.err
  python - <<PY >> $O
import json
try:
    d=json.loads(open('gpurun_out/fm_$fold.json').read().strip().splitlines()[-1])
    ra=d['roofline_agg']
    print('fold $fold rep $rep', 'step %.2f ms'%d['ms_per_step'], 'agg fwd %.3f ms (%.3f)'%(ra['fwd']['avg_ms'],ra['fwd']['frac']), 'bwd %.3f ms (%.3f)'%(ra['bwd']['avg_ms'],ra['bwd']['frac']), 'sm_mhz', d['clocks'].get('sm_mhz'), d['kernel_time_shares'])
except Exception as e:
    print('fold $fold failed', e, open('gpurun_out/fm_$fold.err').read()[-800:])
PY
done; done
cat $O
