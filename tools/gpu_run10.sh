#!/bin/bash
# 2-GPU: multi-GPU parity check, then bench at N=2; plus single-GPU band sweep on rank 0's GPU
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py > gpurun_out/multi_gpu_check.log 2>&1; echo "check rc=$?" | tee gpurun_out/summary.txt
grep -E "rank 0|MULTI" gpurun_out/multi_gpu_check.log | tail -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench_n2.json; grep -E "ms/epoch|Error|error" gpurun_out/bench_n2.err | tail -5
for mb in 0 16 32 64; do
BAND_MB=$mb timeout 300 python tools/band_sweep.py >> gpurun_out/band_sweep.log 2>&1
done
cat gpurun_out/band_sweep.log | grep -v "^data"
