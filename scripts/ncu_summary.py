"""Condense an `ncu --set full` report into the per-launch CSV kept under profiles/.

usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep profiles/rN_name_summary.csv
"""
import csv
import io
import subprocess
import sys

KEEP = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
    "smsp__cycles_active.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    idx = [hdr.index(k) for k in KEEP if k in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        for r in rows:
            w.writerow([r[i] for i in idx])
    print(f"{out}: {len(rows) - 2} launches, {len(idx)} metrics")


if __name__ == "__main__":
    main()
