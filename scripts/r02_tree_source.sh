#!/bin/bash
# Source-level ncu capture of the per-tree aggregation kernels of one training step at the bench size (HEAD):
# per-launch metrics, stall reasons and executed instructions per CUDA source line (scripts/ncu_source_lines.py).
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gat_tree_bwd_kernel|gat_tree_fwd_kernel' \
  --launch-skip 36 --launch-count 12 -o /tmp/tree -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-small --stream-steps 0 > gpurun_out/tree_ncu.log 2>&1
ls -la /tmp/tree.ncu-rep
# launch ids are 1-based in capture order: forward launches come first, then the backward ones
python scripts/ncu_source_lines.py /tmp/tree.ncu-rep 12 45 > gpurun_out/r02_ncu_tree_bwd_source.txt 2>&1
python scripts/ncu_source_lines.py /tmp/tree.ncu-rep 1 30 > gpurun_out/r02_ncu_tree_fwd_source.txt 2>&1
head -60 gpurun_out/r02_ncu_tree_bwd_source.txt
