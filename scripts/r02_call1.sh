#!/bin/bash
# round 2, call 1: bisection of the in-degree > 4 gradient deviation, suite + bench as they stand
mkdir -p gpurun_out
timeout 300 python scripts/deg5_bisect.py > gpurun_out/r02_deg5_bisect.txt 2>&1; echo "rc=$?" >> gpurun_out/r02_deg5_bisect.txt
timeout 300 python scripts/deg5_check.py > gpurun_out/r02_deg5_check.txt 2>&1; echo "rc=$?" >> gpurun_out/r02_deg5_check.txt
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_start.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_start.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_start.json 2> gpurun_out/r02_bench_start.err
tail -n 80 gpurun_out/r02_deg5_bisect.txt
