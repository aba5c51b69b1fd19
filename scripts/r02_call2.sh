#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r02_pytest_2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_2.log
tail -n 40 gpurun_out/r02_pytest_2.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r02_smoke_2.log 2>&1; tail -n 5 gpurun_out/r02_smoke_2.log
