#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "ccdpp or cli" > gpurun_out/pytest_sub.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/summary.txt
tail -3 gpurun_out/pytest_sub.log
timeout 300 python tools/bench_solvers.py --algo ccdpp --rank 64 > gpurun_out/solver_ccdpp.json 2> gpurun_out/solver_ccdpp.err; echo "ccdpp rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/solver_ccdpp.json
C="python tools/bench_solvers.py --algo ccdpp --rank 64 --scale 0.2"
$C > gpurun_out/plain_ccd.json 2> gpurun_out/plain_ccd.err &&
ncu --set full --clock-control none --import-source on -k regex:ccd_window -s 100 -c 4 -f -o gpurun_out/prof_ccd_r1 $C > gpurun_out/ncu_ccd.log 2>&1
echo "ccd ncu rc=$?" | tee -a gpurun_out/summary.txt
