#!/bin/bash
mkdir -p gpurun_out
health() { timeout 60 nvidia-smi --query-gpu=index,memory.used --format=csv,noheader || { echo "GPU UNHEALTHY after $1"; exit 7; }; }
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $T --master-port 29541 bench.py --gpus 2 --steps 30 --warmup 5 --cuda-graphs 1 --profile-host > gpurun_out/bench2_g1_full.log 2>&1; health g1
grep "^{" gpurun_out/bench2_g1_full.log | tail -1 > gpurun_out/bench2_g1.json
python - gpurun_out/bench2_g1.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d["value"]), d["ms_per_step"], "host_enqueue_ms", round(d.get("host_enqueue_ms_per_step"),3), "launches", d["gpu_launches"], "e2e", round(d["e2e"]["value"]), d["e2e"]["ms_per_step"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
grep -v "^$" gpurun_out/bench2_g1_full.log | grep -v "^\[rank1\]" | grep -A 52 "function calls" | cut -c1-200 | head -70
grep -n "Error" gpurun_out/bench2_g1_full.log | head -5
