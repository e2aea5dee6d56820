#!/bin/bash
# multi-GPU measurements on one box: DSGD bench at N = 8, 4, 2 and row-sharded ALS / CCD++ at N = 8
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l); echo "gpus $NG" | tee gpurun_out/summary.txt
run() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) "$@"; }
for n in 8 4 2; do
  [ $n -le $NG ] || continue
  run $n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; echo "bench n$n rc=$?" | tee -a gpurun_out/summary.txt
  python -c "import json;d=json.load(open('gpurun_out/bench_n$n.json'));print($n,d['value']/1e9,'G/s',d['ms_per_step'],'ms','rmse',d['val_rmse'],d['roofline'].get('per_rank_nnz'))"
done
n=$NG
run $n tools/bench_solvers.py --algo als --rank 128 > gpurun_out/solver_als_n$n.json 2> gpurun_out/solver_als_n$n.err; echo "als n$n rc=$?" | tee -a gpurun_out/summary.txt; cat gpurun_out/solver_als_n$n.json
run $n tools/bench_solvers.py --algo ccdpp --rank 64 > gpurun_out/solver_ccdpp_n$n.json 2> gpurun_out/solver_ccdpp_n$n.err; echo "ccdpp n$n rc=$?" | tee -a gpurun_out/summary.txt; cat gpurun_out/solver_ccdpp_n$n.json
