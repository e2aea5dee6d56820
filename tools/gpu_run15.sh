#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "ccdpp or eval" > gpurun_out/pytest_ccd.log 2>&1; echo "pytest ccd rc=$?" | tee gpurun_out/summary.txt
tail -3 gpurun_out/pytest_ccd.log
timeout 300 python tools/bench_solvers.py --algo ccdpp --rank 64 > gpurun_out/solver_ccdpp.json 2> gpurun_out/solver_ccdpp.err; echo "ccdpp rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/solver_ccdpp.json
C="python tools/bench_solvers.py --algo ccdpp --rank 64 --scale 0.2"
$C > gpurun_out/plain_ccd.json 2> gpurun_out/plain_ccd.err &&
ncu --set full --clock-control none --import-source on -k regex:ccd_ -s 200 -c 6 -f -o gpurun_out/prof_ccd_r1 $C > gpurun_out/ncu_ccd.log 2>&1
echo "ccd ncu rc=$?" | tee -a gpurun_out/summary.txt
