#!/bin/bash
# profile captures for profiles/: launch list of the bench command, full captures of the top kernels
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1"
$B > gpurun_out/plain_bench.json 2> gpurun_out/plain_bench.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_r1.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?" | tee gpurun_out/summary.txt
ncu --set full --clock-control none --import-source on -k regex:sgd_flat -s 1 -c 1 -f -o gpurun_out/prof_sgd_flat_r1 $B > gpurun_out/ncu_sgd_flat.log 2>&1
echo "sgd_flat rc=$?" | tee -a gpurun_out/summary.txt
ncu --set full --clock-control none --import-source on -k regex:sgd_run -s 8 -c 1 -f -o gpurun_out/prof_sgd_run_r1 $B > gpurun_out/ncu_sgd_run.log 2>&1
echo "sgd_run rc=$?" | tee -a gpurun_out/summary.txt
A="python tools/bench_solvers.py --algo als --rank 128 --scale 0.1 --epochs 1"
$A > gpurun_out/plain_als.json 2> gpurun_out/plain_als.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:als_ --csv --log-file gpurun_out/launches_als_r1.csv $A > gpurun_out/ncu_als_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:als_gram_tc -s 2 -c 2 -f -o gpurun_out/prof_als_tc_r1 $A > gpurun_out/ncu_als.log 2>&1
echo "als rc=$?" | tee -a gpurun_out/summary.txt
C="python tools/bench_solvers.py --algo ccdpp --rank 64 --scale 0.2"
$C > gpurun_out/plain_ccd.json 2> gpurun_out/plain_ccd.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ccd_\|col_ --csv --log-file gpurun_out/launches_ccd_r1.csv $C > gpurun_out/ncu_ccd_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ccd_ -s 200 -c 6 -f -o gpurun_out/prof_ccd_r1 $C > gpurun_out/ncu_ccd.log 2>&1
echo "ccd rc=$?" | tee -a gpurun_out/summary.txt
V="python tools/bench_solvers.py --algo eval --rank 64 --scale 0.5"
$V > gpurun_out/plain_eval.json 2> gpurun_out/plain_eval.err &&
ncu --set full --clock-control none --import-source on -k regex:eval_sse -s 3 -c 1 -f -o gpurun_out/prof_eval_r1 $V > gpurun_out/ncu_eval.log 2>&1
echo "eval rc=$?" | tee -a gpurun_out/summary.txt
ls -la gpurun_out/*.ncu-rep
