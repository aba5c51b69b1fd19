#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -k "runners" 2>&1 | tail -3
timeout 600 python scripts/e2e_timeline.py 2>&1 | tee gpurun_out/r02_e2e_timeline.txt | tail -12
