#!/bin/bash
# full ncu captures of the key-switch kernels (giant-step batches); only text exports are kept (the .ncu-rep files of these
# heavily unrolled kernels are > 64 MiB with --import-source)
set -u
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
for spec in "k_ks_inner2:6:3" "k_ks_moddown2:6:3" "k_ntt2_inv:8:3"; do
  IFS=: read k s c <<< "$spec"
  timeout 600 ncu --set full --clock-control none -k regex:$k -s $s -c $c -f -o /tmp/prof_$k $B > gpurun_out/prof_$k.log 2>&1
  ncu -i /tmp/prof_$k.ncu-rep --page raw --csv > gpurun_out/prof_$k.raw.csv 2>/dev/null
  ncu -i /tmp/prof_$k.ncu-rep --page details > gpurun_out/prof_$k.details.txt 2>/dev/null
done
ls -la gpurun_out
