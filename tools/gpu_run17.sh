#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/summary.txt
tail -4 gpurun_out/pytest_gpu.log
timeout 400 python tools/bench_solvers.py --algo als --rank 128 --tc 1 > gpurun_out/solver_als_tc1.json 2> gpurun_out/solver_als_tc1.err; echo "als rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/solver_als_tc1.json
