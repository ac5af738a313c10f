#!/bin/bash
# Profiling recipe (B200_PROFILING.md): run under gpurun from the repo root. Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
# 1. every launch of our kernels with its device time (shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 8000 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/launches.log 2>&1
# 2. full captures of the top kernels
for spec in "k_mac:0:2" "k_ks_inner:40:2" "k_ks_moddown:40:2" "k_ntt_smem:80:2"; do
  IFS=: read k s c <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c $c -f -o gpurun_out/prof_$k $B > gpurun_out/prof_$k.log 2>&1
done
ls -la gpurun_out
