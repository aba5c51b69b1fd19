#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r02_pytest_10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_10.log
tail -n 6 gpurun_out/r02_pytest_10.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_10.json 2> gpurun_out/r02_bench_10.err; tail -3 gpurun_out/r02_bench_10.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_10.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value']); print(d['e2e']['value'], d['e2e']['ms_per_step']); print(d['small_batch'])
PY
