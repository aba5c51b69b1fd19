#!/bin/bash
# GPU-box pass: parity tests (stop at first failure), smoke(), one bench line per extra config, ragged headline.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
bash scripts/gpu_configs.sh
timeout 300 python bench.py --ragged --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_ragged.json 2>> gpurun_out/configs.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_ragged.json'))
print('ragged', d['ms_per_step'], d['value'], d['config']['nodes_per_gpu'], {k: (round(v['avg_ms'], 3), round(v['frac'], 3)) for k, v in d['roofline_agg'].items()})
PY
