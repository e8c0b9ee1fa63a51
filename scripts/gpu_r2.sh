#!/bin/sh
# one GPU-box pass of round 2: parity tests, bench lines (driver command + secondary workloads), ncu launch list
# and one `ncu --set full` capture of the two kernels of a step.  TAG names the outputs (gpurun_out/${TAG}_*).
TAG=${TAG:-r2}
mkdir -p gpurun_out
if [ "${TESTS:-1}" = "1" ]; then
  (timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log)
  tail -3 gpurun_out/${TAG}_pytest_gpu.log
fi
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_C3.log 2>&1; echo C3 rc=$?
for w in ${WORKLOADS:-}; do
  timeout 600 python bench.py --workload $w --steps ${WSTEPS:-20} --warmup 5 > gpurun_out/${TAG}_bench_$w.log 2>&1; echo $w rc=$?
done
if [ "${REF:-0}" = "1" ]; then
  timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_reference.log 2>&1; echo ref rc=$?
fi
if [ "${NCU:-0}" = "1" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench.csv \
    python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_under_ncu.log 2>&1; echo launches rc=$?
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'qp_kernel|scan_rt_kernel|scan_kernel' -s 8 -c 6 -f -o gpurun_out/${TAG}_full \
    python scripts/prof_run.py C3 12 > gpurun_out/${TAG}_ncu_full.log 2>&1; echo full rc=$?
fi
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.log")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as e:
        print(f, "no json", e); continue
    print(f.split("bench_")[1][:-4], "value %.3g"%d["value"], "ms/step %.4g"%d["ms_per_step"], "e2e %.3g"%d["e2e"]["value"],
          "kernel_us", {k:(round(v,1) if isinstance(v,float) else v) for k,v in d.get("kernel_us",{}).items() if k!="share"},
          "err", d.get("max_pos_err_vs_ref",{}).get("value"), "cpu %.3g"%d.get("cpu_baseline",{}).get("value",0))
PY
