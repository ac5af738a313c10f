#!/bin/bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.raw.csv gpurun_out/prof_*.details.txt
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --synthetic-inputs --workload mm_32k_x_128k_k15_logN14_otf"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 6000 --csv --log-file gpurun_out/launches_otf14.csv $B > gpurun_out/launches_otf14.log 2>&1
for spec in "k_encode:2:1" "k_img_build:4:2"; do
  IFS=: read k s c <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c $c -f -o gpurun_out/prof14_$k $B > gpurun_out/prof14_$k.log 2>&1
  ncu -i gpurun_out/prof14_$k.ncu-rep --page raw --csv > gpurun_out/prof14_$k.raw.csv 2>/dev/null
done
ls -la gpurun_out | grep -i "prof14\|otf14"
tail -3 gpurun_out/launches_otf14.log
