#!/bin/bash
mkdir -p gpurun_out
P=8 EPOCHS=6 timeout 900 python tools/dsgd_emulate.py > gpurun_out/dsgd_emulate_p8.log 2>&1; echo "emulate rc=$?" | tee gpurun_out/summary.txt
cat gpurun_out/dsgd_emulate_p8.log | tail -6
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1"
$B > gpurun_out/plain_bench.json 2> gpurun_out/plain_bench.err &&
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:mfb:: --csv --log-file gpurun_out/launches_bench_r1.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?" | tee -a gpurun_out/summary.txt
