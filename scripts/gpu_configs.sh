#!/bin/bash
# GPU-box pass: one bench line per BASELINE.json config that is not the headline (configs[0], [2], [3]) plus the
# headline at the reference's own TRAIN_BATCH_SIZE (64 trees).  Device-resident numbers only (no e2e / CPU legs).
mkdir -p gpurun_out
: > gpurun_out/configs.jsonl; : > gpurun_out/configs.err
for w in st_gat_3 st_gat_6 st_gat_6_nr st_gcn_3 st_gin_3 st_sage_3; do
  echo "== $w" >> gpurun_out/configs.err
  timeout 240 python bench.py --workload $w --steps 5 --warmup 3 --no-e2e --no-cpu >> gpurun_out/configs.jsonl 2>> gpurun_out/configs.err
  echo "rc=$?" >> gpurun_out/configs.err
done
echo "== st_pgat_spgnn_3 @ 64 trees" >> gpurun_out/configs.err
timeout 240 python bench.py --trees 64 --steps 20 --warmup 5 --e2e-steps 20 >> gpurun_out/configs.jsonl 2>> gpurun_out/configs.err
echo "rc=$?" >> gpurun_out/configs.err
python - <<'PY'
import json
for l in open('gpurun_out/configs.jsonl'):
    try:
        d = json.loads(l)
    except Exception:
        continue
    r = d.get('roofline') or {}
    print(d['metric'], d['config'].get('trees_per_gpu'), 'ms/step %.2f' % d['ms_per_step'], 'graphs/s %.0f' % d['value'],
          'infer ms %.2f' % d['infer']['ms_per_step'], 'dominant', r.get('kernel'), r.get('frac'), d.get('kernel_time_shares'))
PY
tail -30 gpurun_out/configs.err
