#!/bin/bash
# 2-GPU: NCCL/NVLink sharded tests + bench N=2
mkdir -p gpurun_out
health() { timeout 60 nvidia-smi --query-gpu=index,memory.used --format=csv,noheader | tr '\n' ' '; echo; }
timeout 600 python -m pytest tests/test_sharded_ebc_nccl_gpu.py -x -q 2>&1 | tail -4; health
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 2 --steps 50 --warmup 10 > gpurun_out/bench2_final.log 2>&1; echo "rc=$?"; health
grep "^{" gpurun_out/bench2_final.log | tail -1 > gpurun_out/bench2_final.json
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench2_final.json")); print("bench2", round(d["value"]), d["ms_per_step"], "host", round(d["host_enqueue_ms_per_step"],3), "launches", d["gpu_launches"], "e2e", round(d["e2e"]["value"]), d["e2e"]["ms_per_step"], "e2e host", round(d["e2e"]["host_enqueue_ms_per_step"],3), d["config"].get("cpu_binding"), d["clocks"])
except Exception as e:
    print("bench2 FAILED", e)
PY
grep -n "Error\|Traceback" gpurun_out/bench2_final.log | head -5
