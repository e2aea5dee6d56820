#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "truncated_models or zero_learning" > gpurun_out/pytest_hot.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/summary.txt
tail -5 gpurun_out/pytest_hot.log
HOT=1 ITERS=3 python tools/plan_time.py 2>&1 | tail -2 | tee gpurun_out/plan_time.log
