#!/bin/bash
mkdir -p gpurun_out
for c in 2048 16384 65536; do
ALS_CHUNK=$c timeout 400 python tools/bench_solvers.py --algo als --rank 128 --epochs 2 > gpurun_out/solver_als_chunk$c.json 2> gpurun_out/solver_als_chunk$c.err; python -c "import json;d=json.load(open('gpurun_out/solver_als_chunk$c.json'));print($c,d['ms_user'],d['ms_item'])"
done
