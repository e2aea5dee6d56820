#!/bin/bash
# 8-GPU box: correctness check on 2 ranks, then the DSGD bench at N = 8 and N = 4
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > gpurun_out/gpus.txt
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $T --nproc-per-node 2 --master-port 29511 tools/multi_gpu_check.py > gpurun_out/multi_gpu_check.log 2>&1; echo "check rc=$?" | tee gpurun_out/summary.txt
tail -4 gpurun_out/multi_gpu_check.log
timeout 300 $T --nproc-per-node 8 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "n8 rc=$?" | tee -a gpurun_out/summary.txt
cut -c1-260 gpurun_out/bench_n8.json; grep -h "ms/epoch" gpurun_out/bench_n8.err | tail -2
timeout 300 $T --nproc-per-node 4 --master-port 29513 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "n4 rc=$?" | tee -a gpurun_out/summary.txt
cut -c1-260 gpurun_out/bench_n4.json
timeout 300 $T --nproc-per-node 2 --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "n2 rc=$?" | tee -a gpurun_out/summary.txt
cut -c1-260 gpurun_out/bench_n2.json
