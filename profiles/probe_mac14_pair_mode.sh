#!/bin/bash
set -u
mkdir -p gpurun_out
W=mm_16k_x_64k_k15_logN14
for cfg in "0 0" "0 2" "1 0" "1 2" "0 4" "0 1"; do
  set -- $cfg
  SFG_TC_NOPAIR=$1 SFG_TC_DBG=$2 timeout -s KILL 300 python bench.py --workload $W --steps 2 --warmup 3 --no-cpu-baseline --synthetic-inputs > gpurun_out/b31.json 2> gpurun_out/b31.err
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/b31.json") if l.startswith("{")][-1]
    print("nopair=$1 dbg=$2", "mac kernel ms", round(d["roofline"]["avg_launch_ms"],2), "step", round(d["ms_per_step"],1))
except Exception as e:
    print("ERR", e); print(open("gpurun_out/b31.err").read()[-800:])
PY
done | tee gpurun_out/mac14_dbg.txt
