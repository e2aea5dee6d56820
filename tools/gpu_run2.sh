#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee gpurun_out/summary.txt
timeout 900 python tools/diag_sgd.py > gpurun_out/diag_sgd.log 2>&1; echo "diag rc=$?" | tee -a gpurun_out/summary.txt
timeout 1500 python tools/sweep_sgd.py > gpurun_out/sweep_sgd.log 2>&1; echo "sweep rc=$?" | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/smoke.log
