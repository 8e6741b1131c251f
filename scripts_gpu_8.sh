#!/bin/bash
# 8-GPU box: bench at N=8 and N=4 (device-timed + e2e), health checks in between
mkdir -p gpurun_out
health() { timeout 60 nvidia-smi --query-gpu=index,memory.used --format=csv,noheader | tr '\n' ' '; echo; }
for n in 8; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29550+n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/bench${n}_full.log 2>&1; echo "rc=$?"; health
grep "^{" gpurun_out/bench${n}_full.log | tail -1 > gpurun_out/bench${n}.json
python - gpurun_out/bench${n}.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d["value"]), d["ms_per_step"], "host_enqueue_ms", round(d.get("host_enqueue_ms_per_step"),3), "launches", d["gpu_launches"], "e2e", round(d["e2e"]["value"]), d["e2e"]["ms_per_step"], d["clocks"], d["config"]["parallelism"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
grep -n "Error\|error\|Traceback" gpurun_out/bench${n}_full.log | head -5
done
