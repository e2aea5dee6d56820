#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/flat_limits.py > gpurun_out/flat_limits.log 2>&1; echo "rc=$?"
cat gpurun_out/flat_limits.log
