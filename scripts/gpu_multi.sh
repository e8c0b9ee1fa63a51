#!/bin/sh
# multi-GPU bench lines on one box: sh scripts/gpu_multi.sh <ngpus> "<workloads>"
N=$1; shift
mkdir -p gpurun_out
for w in $1; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --workload $w --steps 20 --warmup 5 > gpurun_out/r2n_bench_${w}_${N}gpu.log 2>&1
  echo "$w x$N rc=$?"; grep '^{' gpurun_out/r2n_bench_${w}_${N}gpu.log | tail -1 | cut -c 1-600
done
