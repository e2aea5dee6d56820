#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ccdpp" > gpurun_out/pytest_ccd.log 2>&1; echo "pytest ccd rc=$?" | tee gpurun_out/summary.txt
tail -3 gpurun_out/pytest_ccd.log
MFB_CCD_FUSE=0 python tools/bench_solvers.py --algo ccdpp --rank 64 > gpurun_out/solver_ccdpp_fuse0.json 2> gpurun_out/solver_ccdpp_fuse0.err; cat gpurun_out/solver_ccdpp_fuse0.json
MFB_CCD_FUSE=1 python tools/bench_solvers.py --algo ccdpp --rank 64 > gpurun_out/solver_ccdpp.json 2> gpurun_out/solver_ccdpp.err; cat gpurun_out/solver_ccdpp.json
