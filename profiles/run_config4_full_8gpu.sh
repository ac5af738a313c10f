#!/bin/bash
set -u
mkdir -p gpurun_out
W=pca_100k_x_500k_k15_logN14_otf
n=8
timeout 840 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29515 bench.py --workload $W --gpus $n --steps 2 --warmup 3 > gpurun_out/config4_full_n$n.json 2> gpurun_out/config4_full_n$n.err
nvidia-smi --query-gpu=memory.used --format=csv,noheader | head -2
python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/config4_full_n$n.json") if l.startswith("{")][-1]
    print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, d["phases_ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "verify", d["verify"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/config4_full_n$n.err").read()[-3000:])
PY
