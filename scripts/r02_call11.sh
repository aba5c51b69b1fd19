#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -8
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-small > gpurun_out/r02_bench_11.json 2> gpurun_out/r02_bench_11.err; tail -2 gpurun_out/r02_bench_11.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_11.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value']); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step']); print('stream', d['stream_1M_trees']['ms_per_step'])
PY
timeout 300 python scripts/e2e_timeline2.py 2>&1 | tail -3
