#!/bin/bash
# Round-2 evidence pass on one B200: parity suite, smoke, both bench arms, launch list, ncu --set full of the top kernels
# (summarised on the box), other BASELINE configs.  Everything lands in gpurun_out/r02_final_*.
mkdir -p gpurun_out
P=gpurun_out/r02_final
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > ${P}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> ${P}_pytest_gpu.log; tail -3 ${P}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1; tail -1 ${P}_smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > ${P}_bench_reference_arm.json 2> ${P}_bench.err
timeout 900 python bench.py --steps 10 --warmup 3 > ${P}_bench.json 2>> ${P}_bench.err; tail -2 ${P}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file ${P}_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-small --stream-steps 0 > ${P}_b_ncu.log 2>&1
python scripts/launch_summary.py ${P}_launches.csv > ${P}_launches_summary.txt 2>&1; head -14 ${P}_launches_summary.txt
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'nt_planes_kernel|nt_pair_kernel|tn_planes_kernel|wide_kernel|gat_tree_fwd_kernel|gat_tree_bwd_kernel|aggx_|split_planes_kernel|reduce_splits' \
  --launch-skip 95 --launch-count 60 -o /tmp/full_step -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-small --stream-steps 0 > ${P}_b_full.log 2>&1
ls -la /tmp/full_step.ncu-rep
python scripts/ncu_traffic.py /tmp/full_step.ncu-rep ${P}_ncu_traffic.json 4096 > ${P}_ncu_table.txt 2>&1; cat ${P}_ncu_table.txt
python scripts/ncu_summary.py /tmp/full_step.ncu-rep 10 > ${P}_ncu_full_summary.txt 2>&1
ncu -i /tmp/full_step.ncu-rep --page raw --csv 2>/dev/null | gzip > ${P}_ncu_raw.csv.gz
if [ -n "$SKIP_CONFIGS" ]; then ls -la gpurun_out | tail -20; exit 0; fi      # other-config lines unchanged since the last pass
: > ${P}_configs.jsonl
for w in st_gat_3 st_gat_6 st_gat_6_nr st_gcn_3 st_gin_3 st_sage_3; do
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-e2e --no-cpu --no-small --stream-steps 0 >> ${P}_configs.jsonl 2>> ${P}_bench.err
done
timeout 300 python bench.py --ragged --steps 5 --warmup 3 --no-e2e --no-cpu --no-small --stream-steps 0 >> ${P}_configs.jsonl 2>> ${P}_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/r02_final_configs.jsonl'):
    try: d = json.loads(l)
    except Exception: continue
    ra = d.get('roofline_agg') or {}
    print(d['metric'], d['config']['workload'][-24:], 'ms/step %.2f' % d['ms_per_step'], 'graphs/s %.0f' % d['value'], 'infer %.0f' % d['infer']['value'],
          'agg fwd %.3f' % ((ra.get('fwd') or {}).get('frac') or 0), 'bwd %.3f' % ((ra.get('bwd') or {}).get('frac') or 0))
PY
ls -la gpurun_out | tail -20
