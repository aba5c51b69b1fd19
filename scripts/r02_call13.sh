#!/bin/bash
# same-box A/B: branched vs branch-free destination side, gather batch 2/3/4, prefetch depth 1/2/3 (640 threads x 4 rounds)
mkdir -p gpurun_out
O=gpurun_out/r02_tree_ab3.txt
: > $O
run() {  # label lib variant
  if [ -n "$2" ]; then export SPGNN_B200_LIB=$PWD/spgnn_b200/$2; else unset SPGNN_B200_LIB; fi
  SPGNN_TREE_BWD=$3 timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-small --stream-steps 0 > gpurun_out/ab3.json 2>gpurun_out/ab3.err
  python - "$1" <<'PY' >> gpurun_out/r02_tree_ab3.txt
import json, sys
try:
    d=json.loads(open('gpurun_out/ab3.json').read().strip().splitlines()[-1])
    ra=d['roofline_agg']
    print('%-22s'%sys.argv[1], 'step %.2f ms'%d['ms_per_step'], 'agg fwd %.3f ms (%.3f)'%(ra['fwd']['avg_ms'],ra['fwd']['frac']), 'bwd %.3f ms (%.3f)'%(ra['bwd']['avg_ms'],ra['bwd']['frac']), 'sm_mhz', d['clocks'].get('sm_mhz'))
except Exception as e:
    print(sys.argv[1], 'failed', e, open('gpurun_out/ab3.err').read()[-600:])
PY
}
for rep in 1 2 3; do
  run "branched kb3 pf2" lib_base.so 3
  run "branchfree kb3 pf2" "" 3
  run "branchfree kb2 pf2" lib_kb2.so 3
  run "branchfree kb4 pf2" lib_kb4.so 3
  run "branchfree kb3 pf1" "" 8
  run "branchfree kb3 pf3" "" 9
done
cat $O
