#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "ccdpp or als or cli or smoke or column" > gpurun_out/pytest_sub.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/summary.txt
tail -4 gpurun_out/pytest_sub.log
timeout 300 python tools/bench_solvers.py --algo ccdpp --rank 64 > gpurun_out/solver_ccdpp.json 2> gpurun_out/solver_ccdpp.err; echo "ccdpp rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/solver_ccdpp.json
timeout 400 python tools/bench_solvers.py --algo als --rank 128 --tc 1 > gpurun_out/solver_als_tc1.json 2> gpurun_out/solver_als_tc1.err; echo "als rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/solver_als_tc1.json
