#!/bin/sh
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest_gpu.log)
tail -3 gpurun_out/r2c_pytest_gpu.log
timeout 300 python scripts/iter_stats.py C3 N2000 C4 > gpurun_out/r2c_iter_stats.log 2>&1; echo iter rc=$?
DMPCB200_SCAN_PRUNE=1024 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 200 --csv --log-file gpurun_out/r2c_launches_pruned_N2000.csv python scripts/scan_probe.py N2000 > gpurun_out/r2c_pruned_under_ncu.log 2>&1; echo ncu rc=$?
