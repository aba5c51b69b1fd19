#!/bin/bash
# GPU-box pass 2: parity tests, ncu launch list of the bench command, one `ncu --set full` capture of a whole training
# step's top kernels at the bench size, summarised ON THE BOX (the .ncu-rep is too large to bring back).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/b_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt 2>&1; head -12 gpurun_out/launches_summary.txt
timeout 1200 ncu --set full --clock-control none -k regex:'nt_planes_kernel|tn_planes_kernel|wide_kernel|gat_tree_fwd_kernel|gat_tree_bwd_kernel|aggx_|split_planes_kernel' \
  --launch-skip 90 --launch-count 45 -o /tmp/full_step -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/b_full.log 2>&1
ls -la /tmp/full_step.ncu-rep
python scripts/ncu_traffic.py /tmp/full_step.ncu-rep gpurun_out/ncu_traffic.json 4096 > gpurun_out/ncu_traffic.txt 2>&1; cat gpurun_out/ncu_traffic.txt
python scripts/ncu_summary.py /tmp/full_step.ncu-rep 14 > gpurun_out/ncu_full_summary.txt 2>&1
ncu -i /tmp/full_step.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/full_raw.csv.gz
ls -la gpurun_out
