#!/bin/bash
# GPU-box pass: standalone GEMM checks (cluster multicast on/off), parity tests, bench.
mkdir -p gpurun_out
echo "== planes_check cluster (default)" > gpurun_out/gemm_check.txt
timeout 240 python scripts/planes_check.py --bench > gpurun_out/gemm_cluster.txt 2>&1; rc=$?
cat gpurun_out/gemm_cluster.txt >> gpurun_out/gemm_check.txt; echo "rc=$rc" >> gpurun_out/gemm_check.txt
if [ $rc -ne 0 ] || ! python - <<'PY'
import re,sys
t=open('gpurun_out/gemm_cluster.txt').read()
m=re.search(r'^worst ([0-9.e+-]+)',t,re.M)
sys.exit(0 if m and float(m.group(1))<1e-4 else 1)
PY
then echo "CLUSTER PATH FAILED -> disabling (SPGNN_NT_CLUSTER=1)" | tee -a gpurun_out/gemm_check.txt; export SPGNN_NT_CLUSTER=1; fi
echo "== planes_check SPGNN_NT_CLUSTER=1 (no clusters)" >> gpurun_out/gemm_check.txt
SPGNN_NT_CLUSTER=1 timeout 240 python scripts/planes_check.py --bench >> gpurun_out/gemm_check.txt 2>&1; echo "rc=$?" >> gpurun_out/gemm_check.txt
echo "== wide_check" >> gpurun_out/gemm_check.txt
timeout 240 python scripts/wide_check.py --big >> gpurun_out/gemm_check.txt 2>&1; echo "rc=$?" >> gpurun_out/gemm_check.txt
cat gpurun_out/gemm_check.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
SPGNN_NT_CLUSTER=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_nocluster.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_nocluster.json
