#!/bin/bash
# multi-GPU measurements on one box (data generated on rank 0 and broadcast): parity check on 2 ranks,
# DSGD bench at N = 8, 4, 2, row-sharded ALS at N = 8, 4, 2 and CCD++ at N = 8
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l); echo "gpus $NG" | tee gpurun_out/summary.txt
run() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) "$@"; }
run 2 tools/multi_gpu_check.py > gpurun_out/multi_gpu_check.log 2>&1; echo "check rc=$?" | tee -a gpurun_out/summary.txt
grep -E "rank 0\]|MULTI" gpurun_out/multi_gpu_check.log | tail -8
for n in 8 4 2; do
  [ $n -le $NG ] || continue
  run $n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; echo "bench n$n rc=$?" | tee -a gpurun_out/summary.txt
  python -c "import json;d=json.load(open('gpurun_out/bench_n$n.json'));print($n,d['value']/1e9,'G/s',d['ms_per_step'],'ms','rmse',d['val_rmse'],d['roofline'].get('per_rank_nnz'))"
done
for n in 8 4 2; do
  [ $n -le $NG ] || continue
  run $n tools/bench_solvers.py --algo als --rank 128 > gpurun_out/solver_als_n$n.json 2> gpurun_out/solver_als_n$n.err; echo "als n$n rc=$?" | tee -a gpurun_out/summary.txt; cat gpurun_out/solver_als_n$n.json
done
n=$NG
run $n tools/bench_solvers.py --algo ccdpp --rank 64 > gpurun_out/solver_ccdpp_n$n.json 2> gpurun_out/solver_ccdpp_n$n.err; echo "ccdpp n$n rc=$?" | tee -a gpurun_out/summary.txt; cat gpurun_out/solver_ccdpp_n$n.json
run $n bench.py --impl reference --gpus $n --steps 1 --warmup 0 > gpurun_out/bench_ref_n$n.json 2> /dev/null; echo "ref arm rc=$?" | tee -a gpurun_out/summary.txt
for u in 0 1; do USTORE=$u BAND_MB=0 timeout 300 python tools/band_sweep.py >> gpurun_out/ustore_sweep.log 2>&1; done
grep ustore gpurun_out/ustore_sweep.log
