#!/bin/bash
# One GPU-box pass: parity tests, bench, launch list, full ncu captures of the top kernels (small batch for replay cost).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/b_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'nt_planes_kernel|tn_planes_kernel|wide_kernel|gat_tree_fwd_kernel|gat_tree_bwd_kernel|gat_layer_fwd_kernel|gat_layer_bwd' \
  --launch-skip 60 --launch-count 40 -o gpurun_out/full_top -f python bench.py --trees 512 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/b_full.log 2>&1
ls -la gpurun_out
