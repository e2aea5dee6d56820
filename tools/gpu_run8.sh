#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/summary.txt
tail -5 gpurun_out/pytest_gpu.log
timeout 180 python tools/test_als_tc.py > gpurun_out/als_tc.log 2>&1; echo "als_tc rc=$?" | tee -a gpurun_out/summary.txt
tail -4 gpurun_out/als_tc.log
for tc in 1 0; do
timeout 400 python tools/bench_solvers.py --algo als --rank 128 --tc $tc > gpurun_out/solver_als_tc$tc.json 2> gpurun_out/solver_als_tc$tc.err; echo "als tc=$tc rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/solver_als_tc$tc.json; tail -2 gpurun_out/solver_als_tc$tc.err
done
timeout 300 python tools/bench_solvers.py --algo ccdpp --rank 64 > gpurun_out/solver_ccdpp.json 2> gpurun_out/solver_ccdpp.err; echo "ccdpp rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/solver_ccdpp.json
timeout 300 python tools/bench_solvers.py --algo eval --rank 64 > gpurun_out/solver_eval.json 2> gpurun_out/solver_eval.err; echo "eval rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/solver_eval.json
