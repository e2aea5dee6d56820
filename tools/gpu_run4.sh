#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/summary.txt
tail -15 gpurun_out/pytest_gpu.log | cut -c1-300
B="python bench.py --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 1"
timeout 600 $B > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv $B > gpurun_out/ncu1.log 2>&1
echo "ncu list rc=$?" | tee -a gpurun_out/summary.txt
timeout 600 $B > gpurun_out/bench_small2.json 2> gpurun_out/bench_small2.err && \
ncu --set full --clock-control none --import-source on -k regex:sgd_flat -s 1 -c 2 -o gpurun_out/prof_sgd_flat_r1 $B > gpurun_out/ncu2.log 2>&1
echo "ncu full rc=$?" | tee -a gpurun_out/summary.txt
ls -la gpurun_out/ | tail -8
