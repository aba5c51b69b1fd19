#!/bin/bash
# 2-GPU pass: NCCL data-parallel step == single-process step, then the bench at N=2 (both arms).
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
if [ "$1" != "bench-only" ]; then
timeout 300 $TR --master-port 29511 scripts/dist_check.py > gpurun_out/dist_check.txt 2>&1; echo "rc=$?" >> gpurun_out/dist_check.txt
grep -E "rank|DIST_CHECK|rc=|Error|error" gpurun_out/dist_check.txt | tail -12
fi
timeout 240 $TR --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?"
grep '^{' gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
