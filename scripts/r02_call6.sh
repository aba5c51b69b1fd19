#!/bin/bash
mkdir -p gpurun_out
python scripts/fullsize_precision.py 2>&1 | head -8 | tail -4
for r in 0 8192 4096; do SPGNN_TN_FLUSH_ROWS=$r timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/r02_bench_slabsB$r.json 2>/dev/null; python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_slabsB$r.json').read().strip().splitlines()[-1])
print('flush $r', d['ms_per_step'], d['value'], d['kernel_time_shares'])
PY
done
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "fullsize or gradients_at or linear_fwd_bwd" 2>&1 | tail -3
