#!/bin/bash
# the driver's round-end sequence on one GPU: smoke, default bench, reference arm
mkdir -p gpurun_out
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee gpurun_out/summary.txt
tail -3 gpurun_out/smoke.log
( time python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err ) 2> gpurun_out/bench_r1.time; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench_r1.time | tail -3; cut -c1-250 gpurun_out/bench_r1.json
( time python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r1.json 2> gpurun_out/bench_ref_r1.err ) 2> gpurun_out/bench_ref_r1.time; echo "ref rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench_ref_r1.time | tail -3; cut -c1-400 gpurun_out/bench_ref_r1.json
