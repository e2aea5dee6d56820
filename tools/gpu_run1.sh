#!/bin/bash
# first GPU pass: smoke, parity tests, SGD diagnostics, bench
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/summary.txt
tail -40 gpurun_out/pytest_gpu.log
timeout 900 python tools/diag_sgd.py > gpurun_out/diag_sgd.log 2>&1; echo "diag rc=$?" | tee -a gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
tail -5 gpurun_out/bench_r1.err; cat gpurun_out/bench_r1.json
