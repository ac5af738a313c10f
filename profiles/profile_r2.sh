#!/bin/bash
# Round-2 profiling pass (run under gpurun from the repo root; outputs in gpurun_out/).  B200_PROFILING.md recipe:
#   1. every launch of our kernels with its device time (cold-cache, serialised: compare SHARES, not absolutes) for the default step,
#      the transposed orientation and the on-the-fly (encode inside the call) path
#   2. ncu --set full captures of the top kernels
set -u
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.raw.csv gpurun_out/prof_*.details.txt
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --synthetic-inputs"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 12000 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/launches.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 12000 --csv --log-file gpurun_out/launches_T.csv $B --workload mm_100k_x_10k_k10_logN13_T > gpurun_out/launches_T.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 4000 --csv --log-file gpurun_out/launches_otf13.csv $B --workload mm_10k_x_100k_k10_logN13_otf > gpurun_out/launches_otf13.log 2>&1
for spec in "k_mac_tc:1:1" "k_ks_inner2:6:3" "k_md_accum:0:2"; do
  IFS=: read k s c <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c $c -f -o /tmp/prof_$k $B > gpurun_out/prof_$k.log 2>&1
  ncu -i /tmp/prof_$k.ncu-rep --page raw --csv > gpurun_out/prof_$k.raw.csv 2>/dev/null
  ncu -i /tmp/prof_$k.ncu-rep --page details > gpurun_out/prof_$k.details.txt 2>/dev/null
done
for spec in "k_encode:2:1" "k_img_build:2:2"; do
  IFS=: read k s c <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c $c -f -o /tmp/prof_$k $B --workload mm_10k_x_100k_k10_logN13_otf > gpurun_out/prof_$k.log 2>&1
  ncu -i /tmp/prof_$k.ncu-rep --page raw --csv > gpurun_out/prof_$k.raw.csv 2>/dev/null
  ncu -i /tmp/prof_$k.ncu-rep --page details > gpurun_out/prof_$k.details.txt 2>/dev/null
done
timeout 600 ncu --set full --clock-control none -k regex:k_mac_tc -s 1 -c 1 -f -o /tmp/prof_k_mac_tc_T $B --workload mm_100k_x_10k_k10_logN13_T > gpurun_out/prof_k_mac_tc_T.log 2>&1
ncu -i /tmp/prof_k_mac_tc_T.ncu-rep --page raw --csv > gpurun_out/prof_k_mac_tc_T.raw.csv 2>/dev/null
ls gpurun_out | head -50
