#!/bin/bash
mkdir -p gpurun_out
python scripts/fullsize_precision.py 2>&1 | head -8 | tee gpurun_out/r02_fullsize_precision_slabs8192.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r02_pytest_5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_5.log
tail -n 8 gpurun_out/r02_pytest_5.log
for r in 0 8192 4096; do SPGNN_TN_FLUSH_ROWS=$r timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/r02_bench_slabs$r.json 2>/dev/null; python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_slabs$r.json').read().strip().splitlines()[-1])
print('flush $r', d['ms_per_step'], d['value'], d['kernel_time_shares'])
PY
done
