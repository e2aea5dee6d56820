#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/summary.txt
grep -E "^E  .*AssertionError|^FAILED|passed|failed" gpurun_out/pytest_gpu.log | cut -c1-300
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/bench_r1.err; cat gpurun_out/bench_r1.json | cut -c1-1500
