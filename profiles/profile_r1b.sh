#!/bin/bash
# Profiling pass (run under gpurun from the repo root; outputs in gpurun_out/).  B200_PROFILING.md recipe:
#   1. every launch of our kernels with its device time (cold-cache, serialised: compare SHARES, not absolutes)
#   2. ncu --set full captures of the top kernels; text/csv exports only for the heavily unrolled key-switch kernels
set -u
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.raw.csv gpurun_out/prof_*.details.txt
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 12000 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/launches.log 2>&1
for spec in "k_mac_tc:1:2" "k_ks_inner2:6:3" "k_md_accum:0:2" "k_ntt2_inv:4:4" "k_ks_macd:0:1" "k_ks_moddown2:0:3" "k_md_final:0:3" "k_img_build:2:2"; do
  IFS=: read k s c <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c $c -f -o /tmp/prof_$k $B > gpurun_out/prof_$k.log 2>&1
  ncu -i /tmp/prof_$k.ncu-rep --page raw --csv > gpurun_out/prof_$k.raw.csv 2>/dev/null
  ncu -i /tmp/prof_$k.ncu-rep --page details > gpurun_out/prof_$k.details.txt 2>/dev/null
done
cp /tmp/prof_k_mac_tc.ncu-rep gpurun_out/ 2>/dev/null
ls -la gpurun_out
