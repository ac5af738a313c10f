#!/bin/bash
# Stress for the round-1 flake: N cycles of "a process that loads the GPU heavily" -> "a FRESH process whose first
# Preprocess + Compute (and first ciphertext-algebra calls) are compared bit for bit with the oracle".  Every mismatch is a hard test
# failure (the arbiter prints which side moved).  Usage (under gpurun, repo root): bash profiles/stress_parity.sh [cycles] [poison]
set -u
N=${1:-20}
POISON=${2:-0}
mkdir -p gpurun_out
LOG=gpurun_out/stress_parity.log
: > $LOG
fail=0
for i in $(seq 1 $N); do
  python bench.py --workload mm_2k_x_20k_k10_logN13 --steps 2 --warmup 3 --no-cpu-baseline --synthetic-inputs > /dev/null 2>> $LOG
  SFG_POISON=$POISON python -m pytest tests/test_gpu_parity.py tests/test_gpu_ctalg.py -q -m gpu -p no:cacheprovider \
      -k "preprocess_compute_bit_exact or mul_relin_rescale or qx_lazy or pn14_shape" > /tmp/stress_$i.log 2>&1
  rc=$?
  tail -1 /tmp/stress_$i.log | sed "s/^/cycle $i (poison=$POISON): /" >> $LOG
  if [ $rc -ne 0 ]; then fail=$((fail+1)); cat /tmp/stress_$i.log >> $LOG; fi
done
echo "stress_parity: $N cycles, $fail with failures (poison=$POISON)" | tee -a $LOG
exit $fail
