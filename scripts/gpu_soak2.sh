#!/bin/sh
# wider soak: other seeds, the other variants, every scenario of the C5 batch
mkdir -p gpurun_out
{
  echo "# wider soak vs the oracle: seeds / variants / C5 scenarios.  $(date -u +%Y-%m-%dT%H:%MZ), $(nvidia-smi --query-gpu=name --format=csv,noheader | head -1)"
  timeout 1500 python scripts/soak_vs_oracle.py C5 149
  for seed in 1 2 3 4 5 6; do timeout 300 python scripts/soak_vs_oracle.py C3:$seed 60; done
  for seed in 7 8; do timeout 300 python scripts/soak_vs_oracle.py N2000:$seed 30; done
  for seed in 9 10; do timeout 300 python scripts/soak_vs_oracle.py C4:$seed 25; done
  for v in 1 3; do for seed in 11 21 22 23 31 32 33 34; do timeout 300 python scripts/soak_vs_oracle.py C3:$seed:$v 60; done; timeout 300 python scripts/soak_vs_oracle.py N100:12:$v 100; timeout 300 python scripts/soak_vs_oracle.py N2000:24:$v 20; done
  for seed in 13 14 15; do timeout 300 python scripts/soak_vs_oracle.py C2:$seed 60; done
} 2>&1 | tee gpurun_out/r2k_soak_wide.txt
