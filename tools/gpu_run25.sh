#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "sgd or smoke or cli_sgd or trains_from_memory" > gpurun_out/pytest_sgd.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/summary.txt
tail -12 gpurun_out/pytest_sgd.log
timeout 600 python tools/flat_limits.py > gpurun_out/flat_limits2.log 2>&1; cat gpurun_out/flat_limits2.log
timeout 900 python bench.py --no-solvers --no-cpu-baseline > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
python -c "import json;d=json.load(open('gpurun_out/bench_r1c.json'));print(d['value'],d['ms_per_step'],d['val_rmse'],d['e2e'])"
