#!/bin/sh
# one GPU-box pass: parity tests, then the bench on the headline and the secondary workloads
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log)
tail -3 gpurun_out/pytest_gpu.log
show() {
  tail -1 "$1" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step']*1e3,1), 'us/step; graph', d['resident_graph']['value'] and round(d['resident_graph']['value']), 'e2e', round(d['e2e']['value']), 'kernels', {k: round(v,1) for k,v in d['kernel_us'].items() if k!='share'})
"
}
for w in ${WORKLOADS:-C3 C4}; do
  timeout 400 python bench.py --workload $w --steps 60 --warmup 4 > gpurun_out/bench_$w.log 2>&1; echo $w rc=$?; show gpurun_out/bench_$w.log
done
for b in ${BIGS:-0}; do
  DMPCB200_SCAN_BIG=$b timeout 400 python bench.py --workload N2000 --steps 60 --warmup 4 > gpurun_out/bench_N2000_$b.log 2>&1; echo N2000 big=$b rc=$?; show gpurun_out/bench_N2000_$b.log
done
