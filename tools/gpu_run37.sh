#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/summary.txt
tail -3 gpurun_out/pytest_gpu.log
HOT=1,0 ITERS=4 python tools/plan_time.py 2>&1 | tail -8 | tee gpurun_out/plan_time.log
