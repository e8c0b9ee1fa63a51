#!/bin/sh
# whole transitions teacher-forced against the oracle (scripts/soak_vs_oracle.py); output kept under profiles/
mkdir -p gpurun_out
{
  echo "# soak: every step of a closed-loop transition on the GPU against one oracle step on the same inputs"
  echo "# (status flags, retry counts, horizons; tolerance 1e-6 m).  $(date -u +%Y-%m-%dT%H:%MZ), $(nvidia-smi --query-gpu=name --format=csv,noheader | head -1)"
  timeout 900 python scripts/soak_vs_oracle.py C3 149
  timeout 300 python scripts/soak_vs_oracle.py N100 149
  timeout 600 python scripts/soak_vs_oracle.py C2 99
  timeout 900 python scripts/soak_vs_oracle.py N2000 100
  timeout 900 python scripts/soak_vs_oracle.py C4 80
} 2>&1 | tee gpurun_out/r2k_soak_vs_oracle.txt
