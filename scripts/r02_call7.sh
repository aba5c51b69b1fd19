#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r02_pytest_7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_7.log
tail -n 8 gpurun_out/r02_pytest_7.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_packed.json 2> gpurun_out/r02_bench_packed.err; tail -3 gpurun_out/r02_bench_packed.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_packed.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value']); print(d['e2e']); print(d['e2e_dense_format']); print(d['h2d_only'])
PY
