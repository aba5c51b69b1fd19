#!/bin/bash
# CTA pairs in the NT and wide GEMMs: parity suite, wide check, A/B of the headline step
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/r02_pytest_15.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_15.log
timeout -k 5 200 python scripts/wide_check.py > gpurun_out/r02_wide_check_pair.txt 2>&1; echo "wide_check rc=$?"; tail -12 gpurun_out/r02_wide_check_pair.txt
O=gpurun_out/r02_pair_ab.txt
: > $O
for rep in 1 2; do
for cfg in "0 0" "1 0" "1 1"; do
  set -- $cfg
  SPGNN_NT_PAIR=$1 SPGNN_WIDE_PAIR=$2 timeout -k 5 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-small --stream-steps 0 > gpurun_out/pair.json 2>gpurun_out/pair.err
  python - "$cfg" <<'PY' >> gpurun_out/r02_pair_ab.txt
import json, sys
try:
    d=json.loads(open('gpurun_out/pair.json').read().strip().splitlines()[-1])
    ra=d['roofline_agg']
    print('nt_pair wide_pair = %s'%sys.argv[1], 'step %.2f ms'%d['ms_per_step'], 'infer %.2f ms'%d['infer']['ms_per_step'], 'sm_mhz', d['clocks'].get('sm_mhz'), d['kernel_time_shares'])
except Exception as e:
    print(sys.argv[1], 'failed', e, open('gpurun_out/pair.err').read()[-600:])
PY
done; done
cat $O
