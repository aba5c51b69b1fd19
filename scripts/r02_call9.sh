#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -k "runners or pe_ or fullsize" 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_packed2.json 2> gpurun_out/r02_bench_packed2.err; tail -3 gpurun_out/r02_bench_packed2.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_packed2.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value']); print(d['e2e']); print(d['e2e_dense_format']); print(d['h2d_only'])
PY
