#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/diag_als.py > gpurun_out/diag_als.log 2>&1; echo "diag_als rc=$?" | tee gpurun_out/summary.txt
cat gpurun_out/diag_als.log
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 -k "sgd or netflix or cli_sgd" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/summary.txt
grep -E "^E  .*AssertionError|^FAILED|passed|failed" gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python tools/diag_sgd.py > gpurun_out/diag_sgd.log 2>&1
