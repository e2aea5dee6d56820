#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; echo "bench rc=$?" | tee gpurun_out/summary.txt
python -c "import json;d=json.load(open('gpurun_out/bench_r1.json'));print(d['value'],d['ms_per_step'],d['val_rmse'],d['clocks']);print(d['e2e']);print(d['cpu_baseline']);print(d.get('solvers'))"
