#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/summary.txt
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python tools/hot_probe.py > gpurun_out/hot_probe2.log 2>&1; tail -8 gpurun_out/hot_probe2.log
B="python bench.py --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 2 --no-solvers"
$B > gpurun_out/plain_bench.json 2> gpurun_out/plain_bench.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/plain_bench.json | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:mfb:: --csv --log-file gpurun_out/launches_bench_r1.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?" | tee -a gpurun_out/summary.txt
ncu --set full --clock-control none --import-source on -k regex:sgd_flat -s 1 -c 1 -f -o gpurun_out/prof_sgd_flat_r1 $B > gpurun_out/ncu_sgd_flat.log 2>&1
echo "sgd_flat ncu rc=$?" | tee -a gpurun_out/summary.txt
ncu --set full --clock-control none --import-source on -k regex:sgd_hot_kernel -s 1 -c 1 -f -o gpurun_out/prof_sgd_hot_r1 $B > gpurun_out/ncu_sgd_hot.log 2>&1
echo "sgd_hot ncu rc=$?" | tee -a gpurun_out/summary.txt
