#!/bin/bash
# round-end sequence on one GPU: full parity suite, smoke, default bench, reference arm, launch list of the bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/summary.txt
tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt
python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
cut -c1-200 gpurun_out/bench_r1.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r1.json 2> gpurun_out/bench_ref_r1.err; echo "ref rc=$?" | tee -a gpurun_out/summary.txt
B="python bench.py --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 2 --no-solvers"
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled --csv --log-file gpurun_out/launches_bench_r1.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?" | tee -a gpurun_out/summary.txt
