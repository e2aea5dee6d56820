#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "als or ccdpp" > gpurun_out/pytest_als.log 2>&1; echo "pytest als/ccd rc=$?" | tee gpurun_out/summary.txt
tail -15 gpurun_out/pytest_als.log
python __graft_entry__.py smoke 2>&1 | tail -2
for d in 1 0; do
ALS_DUAL=$d timeout 400 python tools/bench_solvers.py --algo als --rank 128 --tc 1 > gpurun_out/solver_als_dual$d.json 2> gpurun_out/solver_als_dual$d.err; echo "als dual=$d rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/solver_als_dual$d.json; tail -2 gpurun_out/solver_als_dual$d.err
done
timeout 400 python tools/bench_solvers.py --algo als --rank 64 > gpurun_out/solver_als_r64.json 2> gpurun_out/solver_als_r64.err; cat gpurun_out/solver_als_r64.json
timeout 300 python tools/bench_solvers.py --algo ccdpp --rank 64 > gpurun_out/solver_ccdpp.json 2> gpurun_out/solver_ccdpp.err; echo "ccdpp rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/solver_ccdpp.json
