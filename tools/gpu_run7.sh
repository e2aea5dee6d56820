#!/bin/bash
mkdir -p gpurun_out
timeout 180 python tools/test_als_tc.py > gpurun_out/als_tc.log 2>&1; echo "als_tc rc=$?" | tee gpurun_out/summary.txt
cat gpurun_out/als_tc.log | tail -8
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
tail -2 gpurun_out/bench_r1.err
