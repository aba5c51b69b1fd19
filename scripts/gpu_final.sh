#!/bin/bash
# Round-end style pass on one GPU: parity tests, smoke(), both bench arms, ncu launch list of the bench command.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench.err; cat gpurun_out/bench_ref.json
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2>> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/b_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt 2>&1; head -30 gpurun_out/launches_summary.txt
