#!/bin/bash
# One parameterised runner for a gpurun call: runs the named stages, every stage under its own timeout, logs under
# gpurun_out/<tag>_*.  usage: tools/gpu_session.sh <tag> <stage> [<stage> ...]
#   stages: tests | tests:<pytest -k expr> | l2 | curves | bench | bench:<extra args> | ref | ncu_launches | cmd:<shell command>
set -u
tag=$1; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
for st in "$@"; do
  name=${st%%:*}; arg=""; [[ "$st" == *:* ]] && arg=${st#*:}
  t0=$(date +%s)
  case $name in
    tests) if [ -n "$arg" ]; then timeout 900 python -m pytest tests -x -q -m gpu --timeout 240 -k "$arg" > gpurun_out/${tag}_tests.log 2>&1; else timeout 900 python -m pytest tests -x -q -m gpu --timeout 240 > gpurun_out/${tag}_tests.log 2>&1; fi; echo "rc=$?" >> gpurun_out/${tag}_tests.log; tail -5 gpurun_out/${tag}_tests.log;;
    l2) timeout 120 tools/l2_probe > gpurun_out/${tag}_l2_probe.json 2> gpurun_out/${tag}_l2_probe.err; cat gpurun_out/${tag}_l2_probe.json;;
    curves) timeout 900 python tools/dsgd_device_curves.py $arg --out gpurun_out/${tag}_dsgd_device_curves.json > gpurun_out/${tag}_curves.log 2>&1; tail -30 gpurun_out/${tag}_curves.log | cut -c1-250;;
    bench) timeout 900 python bench.py $arg > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "rc=$?"; tail -3 gpurun_out/${tag}_bench.err; head -c 1500 gpurun_out/${tag}_bench.json;;
    ref) timeout 600 python bench.py --impl reference $arg > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "rc=$?"; cat gpurun_out/${tag}_bench_ref.json;;
    cmd) timeout 600 bash -c "$arg" > gpurun_out/${tag}_cmd.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/${tag}_cmd.log;;
  esac
  echo "== stage $st took $(( $(date +%s) - t0 )) s"
done
