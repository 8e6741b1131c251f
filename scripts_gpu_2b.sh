#!/bin/bash
mkdir -p gpurun_out
health() { timeout 60 nvidia-smi --query-gpu=index,memory.used --format=csv,noheader || { echo "GPU UNHEALTHY after $1"; exit 7; }; }
timeout 400 python -m pytest "tests/test_sharded_ebc_nccl_gpu.py::test_fused_nvlink_path[twrw]" -x -q -s 2>&1 | grep -v "^$" | tail -60 > gpurun_out/g2_twrw.log; grep -n "Fatal\|File \|Error\|passed\|failed" gpurun_out/g2_twrw.log | tail -30; health twrw
TRB_GRAD_PUSH=0 timeout 400 python -m pytest "tests/test_sharded_ebc_nccl_gpu.py::test_fused_nvlink_path[twrw]" -x -q 2>&1 | tail -3; health twrw_pull
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $T --master-port 29523 bench.py --gpus 2 --steps 30 --warmup 5 --no-e2e 2>&1 | tail -1 > gpurun_out/bench2_push.json; health bench2
timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e 2>&1 | tail -1 > gpurun_out/bench1d.json; health bench1
for f in gpurun_out/bench2_push.json gpurun_out/bench1d.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d["value"]), d["ms_per_step"], "host_enqueue_ms", d.get("host_enqueue_ms_per_step"), d["gpu_launches"])
except Exception as e:
    print(sys.argv[1], "FAILED", e, open(sys.argv[1]).read()[-600:])
PY
done
