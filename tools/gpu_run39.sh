#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:ccd_update -s 24 -c 4 -f -o gpurun_out/prof_ccd_r1 python tools/bench_solvers.py --algo ccdpp --rank 64 > gpurun_out/ncu_ccd.log 2>&1
echo "ccd ncu rc=$?"
