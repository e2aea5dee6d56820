#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "als or smoke" > gpurun_out/pytest_als.log 2>&1; echo "pytest als rc=$?" | tee gpurun_out/summary.txt
tail -15 gpurun_out/pytest_als.log
timeout 180 python tools/test_als_tc.py > gpurun_out/als_tc.log 2>&1; tail -4 gpurun_out/als_tc.log
for tc in 1 0; do
timeout 400 python tools/bench_solvers.py --algo als --rank 128 --tc $tc > gpurun_out/solver_als_tc$tc.json 2> gpurun_out/solver_als_tc$tc.err; echo "als tc=$tc rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/solver_als_tc$tc.json; tail -2 gpurun_out/solver_als_tc$tc.err
done
timeout 400 python tools/bench_solvers.py --algo als --rank 64 > gpurun_out/solver_als_r64.json 2> gpurun_out/solver_als_r64.err; cat gpurun_out/solver_als_r64.json
