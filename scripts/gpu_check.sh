#!/bin/bash
# GPU-box pass: parity tests (stop at first failure), then one bench line per extra config.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
bash scripts/gpu_configs.sh
