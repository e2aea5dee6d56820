#!/bin/bash
mkdir -p gpurun_out
A="python tools/bench_solvers.py --algo als --rank 128 --scale 0.1 --epochs 1"
$A > gpurun_out/plain_als.json 2> gpurun_out/plain_als.err &&
ncu --set full --clock-control none --import-source on -k regex:als_gram_tc -s 2 -c 2 -f -o gpurun_out/prof_als_tc_r1 $A > gpurun_out/ncu_als.log 2>&1
echo "als rc=$?" | tee gpurun_out/summary.txt
cat gpurun_out/plain_als.json
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench_r1.json
