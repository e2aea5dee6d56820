#!/bin/bash
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $T --nproc-per-node 2 --master-port 29521 tools/multi_gpu_check.py > gpurun_out/multi_gpu_check.log 2>&1; echo "check rc=$?" | tee gpurun_out/summary.txt
tail -3 gpurun_out/multi_gpu_check.log
timeout 300 $T --nproc-per-node 2 --master-port 29522 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "n2 rc=$?" | tee -a gpurun_out/summary.txt
cut -c1-230 gpurun_out/bench_n2.json; grep -h "ms/epoch" gpurun_out/bench_n2.err | tail -1
timeout 200 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -2
