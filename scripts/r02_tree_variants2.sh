#!/bin/bash
# second A/B pass: balanced launch shapes (608x4, 832x3) and prefetch depth, plus a launch list of the e2e leg
mkdir -p gpurun_out
O=gpurun_out/r02_tree_variants2.txt
: > $O
for pair in "4 4" "5 0" "6 3" "7 4"; do
  set -- $pair; v=$1; fv=$2
  for cfg in "8 2 64 1 1 1" "8 1 128 1 2 0" "16 2 256 1 1 0"; do
    echo "== check bwd variant $v fwd $fv cfg $cfg" >> $O
    SPGNN_TREE_BWD=$v SPGNN_TREE_FWD=$fv timeout 300 python scripts/tree_check.py bwd $cfg 2>&1 | grep "rel err\|Error\|error" >> $O
  done
done
for rep in 1 2; do
for pair in "0 0" "3 3" "4 4" "5 0" "6 0" "7 0"; do
  set -- $pair; v=$1; fv=$2
  SPGNN_TREE_BWD=$v SPGNN_TREE_FWD=$fv timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-small --stream-steps 0 > gpurun_out/tv_$v.json 2>gpurun_out/tv_$v.err
  python - <<PY >> $O
import json
try:
    d=json.loads(open('gpurun_out/tv_$v.json').read().strip().splitlines()[-1])
    ra=d['roofline_agg']
    print('bwd $v fwd $fv rep $rep', 'step %.2f ms'%d['ms_per_step'], 'infer %.2f ms'%d['infer']['ms_per_step'], 'agg fwd %.3f ms (%.3f)'%(ra['fwd']['avg_ms'],ra['fwd']['frac']), 'bwd %.3f ms (%.3f)'%(ra['bwd']['avg_ms'],ra['bwd']['frac']), 'sm_mhz', d['clocks'].get('sm_mhz'))
except Exception as e:
    print('variant $v failed', e, open('gpurun_out/tv_$v.err').read()[-800:])
PY
done; done
cat $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_e2e_launches.csv \
  python bench.py --steps 1 --warmup 1 --e2e-steps 3 --no-cpu --no-small --stream-steps 0 > gpurun_out/r02_e2e_ncu.log 2>&1
python - <<'PY'
import csv, collections, re
rows=[]
with open('gpurun_out/r02_e2e_launches.csv') as f:
    lines=[l for l in f if not l.startswith('==')]
r=csv.DictReader(lines)
tot=collections.OrderedDict()
for x in r:
    try: v=float(x['Metric Value'].replace(',',''))
    except Exception: continue
    u=x.get('Metric Unit','')
    ms = v/1e6 if u in ('ns','nsecond') else (v/1e3 if u in ('us','usecond') else v)
    n=re.sub(r'\(.*','',x['Kernel Name'])[:70]
    t=tot.setdefault(n,[0,0.0]); t[0]+=1; t[1]+=ms
print('kernel totals over the capture (count, total ms)')
for n,(c,t) in sorted(tot.items(), key=lambda kv:-kv[1][1])[:60]:
    print('%9.3f ms x %5d  %s'%(t,c,n))
PY
