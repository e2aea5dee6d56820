#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/summary.txt
tail -15 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --no-solvers > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
python -c "import json;d=json.load(open('gpurun_out/bench_r1b.json'));print(d['value'],d['e2e'])"
