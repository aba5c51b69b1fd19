#!/bin/bash
# A/B of the per-tree layer kernels' launch shapes (SPGNN_TREE_FWD / SPGNN_TREE_BWD) on one box: correctness of every
# shape against the chunk kernels, then the headline step with each.
mkdir -p gpurun_out
O=gpurun_out/r02_tree_variants.txt
: > $O
for v in 0 1 2 3; do
  fv=$v; [ $v = 3 ] && fv=1
  for cfg in "8 2 64 1 1 1" "8 1 128 1 2 1" "16 2 256 1 1 0" "8 2 64 0 1 1"; do
    echo "== check bwd variant $v fwd $fv cfg $cfg" >> $O
    SPGNN_TREE_BWD=$v SPGNN_TREE_FWD=$fv timeout 300 python scripts/tree_check.py bwd $cfg 2>&1 | grep "rel err\|Error\|error" >> $O
  done
done
for rep in 1 2; do
for v in 0 1 2 3; do
  fv=$v; [ $v = 3 ] && fv=1
  SPGNN_TREE_BWD=$v SPGNN_TREE_FWD=$fv timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-small --stream-steps 0 > gpurun_out/tv_$v.json 2>gpurun_out/tv_$v.err
  python - <<PY >> $O
import json
try:
    d=json.loads(open('gpurun_out/tv_$v.json').read().strip().splitlines()[-1])
    ra=d['roofline_agg']
    print('variant $v rep $rep', 'step %.2f ms'%d['ms_per_step'], 'infer %.2f ms'%d['infer']['ms_per_step'], 'agg fwd %.3f ms (%.3f)'%(ra['fwd']['avg_ms'],ra['fwd']['frac']), 'bwd %.3f ms (%.3f)'%(ra['bwd']['avg_ms'],ra['bwd']['frac']), 'sm_mhz', d['clocks'].get('sm_mhz'))
except Exception as e:
    print('variant $v failed', e, open('gpurun_out/tv_$v.err').read()[-800:])
PY
done; done
cat $O
