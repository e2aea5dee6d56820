#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/plan_time.py > gpurun_out/plan_time.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:mfb:: --csv --log-file gpurun_out/launches_plan.csv python tools/plan_time.py > /dev/null 2>&1
cat gpurun_out/plan_time.log | tail -8
