#!/bin/sh
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest_gpu.log)
tail -3 gpurun_out/r2d_pytest_gpu.log
timeout 300 python scripts/iter_stats.py C4 > gpurun_out/r2d_iter_stats.log 2>&1; echo iter rc=$?
cut -c 1-300 gpurun_out/r2d_iter_stats.log
WORKLOADS="C4" TESTS=0 TAG=r2d sh scripts/gpu_r2.sh
