#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "sgd or cli or entry_point or smoke" > gpurun_out/pytest_hot.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/summary.txt
tail -3 gpurun_out/pytest_hot.log
python tools/bench_variants.py > gpurun_out/bench_variants.json 2> gpurun_out/bench_variants.err; tail -4 gpurun_out/bench_variants.err | cut -c1-330
