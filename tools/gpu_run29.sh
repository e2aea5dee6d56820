#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hot or shuffled_block_order or zero_learning" > gpurun_out/pytest_hot.log 2>&1; echo "pytest hot rc=$?" | tee gpurun_out/summary.txt
tail -5 gpurun_out/pytest_hot.log
timeout 600 python tools/hot_probe.py > gpurun_out/hot_probe.log 2>&1; echo "hot_probe rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/hot_probe.log | tail -20
