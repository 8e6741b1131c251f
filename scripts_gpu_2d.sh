#!/bin/bash
# 1-GPU and 2-GPU bench with sync-debug warnings + host profile
mkdir -p gpurun_out
health() { timeout 60 nvidia-smi --query-gpu=index,memory.used --format=csv,noheader || { echo "GPU UNHEALTHY after $1"; exit 7; }; }
summ() {
grep "^{" $1 | tail -1 > $2
python - $2 <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d["value"]), d["ms_per_step"], "host_enqueue_ms", round(d.get("host_enqueue_ms_per_step"),3), "launches", d["gpu_launches"], "e2e", round(d["e2e"]["value"]), d["e2e"]["ms_per_step"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
grep -i "synchroniz" $1 | grep -v "^\[rank1\]" | sort | uniq -c | sort -rn | cut -c1-260 | head -20
}
true
true
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $T --master-port 29541 bench.py --gpus 2 --steps 30 --warmup 5 --cuda-graphs 1 --profile-host > gpurun_out/bench2_sync_full.log 2>&1; health b2
summ gpurun_out/bench2_sync_full.log gpurun_out/bench2_sync.json
grep -v "^$" gpurun_out/bench2_sync_full.log | grep -v "^\[rank1\]" | grep -A 30 "function calls" | cut -c1-200 | head -40
