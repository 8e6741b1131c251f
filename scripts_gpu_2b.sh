#!/bin/bash
mkdir -p gpurun_out
health() { timeout 60 nvidia-smi --query-gpu=index,memory.used --format=csv,noheader || { echo "GPU UNHEALTHY after $1"; exit 7; }; }
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
port=29530
for v in "g1:--cuda-graphs 1" "g0:--cuda-graphs 0" "g1nodp:--cuda-graphs 1 --dp-rows 0"; do
  tag=${v%%:*}; flags=${v#*:}; port=$((port+1))
  timeout 300 $T --master-port $port bench.py --gpus 2 --steps 30 --warmup 5 $flags 2>&1 | tail -1 > gpurun_out/bench2_$tag.json; health bench2_$tag
  python - gpurun_out/bench2_$tag.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d["value"]), d["ms_per_step"], "host_enqueue_ms", round(d.get("host_enqueue_ms_per_step"),3), "launches", d["gpu_launches"], "e2e", round(d["e2e"]["value"]), d["e2e"]["ms_per_step"], d["clocks"]["samples"])
except Exception as e:
    print(sys.argv[1], "FAILED", e, open(sys.argv[1]).read()[-1500:])
PY
done
timeout 400 python -m pytest "tests/test_sharded_ebc_nccl_gpu.py" -x -q 2>&1 | grep -v "^$" | tail -60 > gpurun_out/g2_pytest.log; grep -n "Fatal\|File \"/tmp\|Error\|passed\|failed" gpurun_out/g2_pytest.log | tail -30; health pytest
