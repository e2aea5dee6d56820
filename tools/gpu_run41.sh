#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sgd" > gpurun_out/pytest_hot.log 2>&1; echo "pytest sgd rc=$?" | tee gpurun_out/summary.txt
tail -3 gpurun_out/pytest_hot.log
HOT=1 ITERS=4 python tools/plan_time.py 2>&1 | tail -4 | tee gpurun_out/plan_time.log
