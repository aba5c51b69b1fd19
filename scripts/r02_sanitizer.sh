#!/bin/bash
# compute-sanitizer over a subset of the parity suite at HEAD: memcheck (out-of-bounds / misaligned accesses) on the
# edge-case graphs, the train-mode step, the wire format and the batch builder; racecheck (shared-memory hazards) on
# the per-tree kernels through one full-width train step.  Small graphs only: the tools slow kernels 10-100x.
mkdir -p gpurun_out
O=gpurun_out/r02_sanitizer.txt
: > $O
echo "== memcheck: edge-case graphs, train-mode step, wire format, builder, PE" >> $O
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_gpu_parity.py tests/test_gpu_runners.py -q -x --timeout 550 \
  -k "edge_case_graphs or train_mode_step or packed_wire or batch_builder or anchor_select or distance_pe_tree or dx_gemm or random_trees" \
  > gpurun_out/sanitizer_memcheck.log 2>&1
echo "rc=$?" >> $O
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds|misaligned" gpurun_out/sanitizer_memcheck.log | head -20 >> $O
echo "== racecheck: full-width 2-tree SPGNN-3 train step (smoke)" >> $O
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 \
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_racecheck.log 2>&1
echo "rc=$?" >> $O
grep -E "RACECHECK SUMMARY|hazard|smoke ok|Error|error" gpurun_out/sanitizer_racecheck.log | head -20 >> $O
echo "== racecheck: general-degree chunk kernels, GCN / GIN / SAGE aggregations (edge-case graphs, full-width fwd+bwd)" >> $O
timeout 240 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_gpu_parity.py -q -x --timeout 230 \
  -k "(edge_case_graphs and trifurcations) or (full_width_forward and (gcn or gin or sage) and 2)" > gpurun_out/sanitizer_racecheck2.log 2>&1
echo "rc=$?" >> $O
grep -E "RACECHECK SUMMARY|hazard|passed|failed|Error" gpurun_out/sanitizer_racecheck2.log | head -20 >> $O
cat $O
