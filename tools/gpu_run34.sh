#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sgd" > gpurun_out/pytest_hot.log 2>&1; echo "pytest sgd rc=$?" | tee gpurun_out/summary.txt
tail -3 gpurun_out/pytest_hot.log
timeout 600 python tools/hot_probe.py > gpurun_out/hot_probe2.log 2>&1; tail -12 gpurun_out/hot_probe2.log
python bench.py --steps 5 --warmup 2 --no-cpu-baseline --e2e-steps 2 --no-solvers > gpurun_out/plain_bench.json 2> gpurun_out/plain_bench.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/plain_bench.json | cut -c1-300; tail -2 gpurun_out/plain_bench.err
