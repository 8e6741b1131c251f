#!/bin/bash
# 2-GPU validation of the NVLink paths; every step is followed by a GPU health check and the script stops at the first problem.
mkdir -p gpurun_out
health() { timeout 60 nvidia-smi --query-gpu=index,memory.used,utilization.gpu --format=csv,noheader || { echo "GPU UNHEALTHY after $1"; exit 7; }; }
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_sharded_ebc_nccl_gpu.py tests/test_tbe_gpu.py tests/test_head_gpu.py tests/test_uvm_cache.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/g2_pytest.log; tail -4 gpurun_out/g2_pytest.log; health pytest
timeout 300 $T --master-port 29523 bench.py --gpus 2 --steps 30 --warmup 5 2>&1 | tail -1 > gpurun_out/bench2_push.json; health bench2
TRB_GRAD_PUSH=0 timeout 300 $T --master-port 29524 bench.py --gpus 2 --steps 30 --warmup 5 --no-e2e 2>&1 | tail -1 > gpurun_out/bench2_pull.json; health bench2pull
timeout 300 $T --master-port 29525 bench.py --gpus 2 --steps 30 --warmup 5 --no-e2e --dp-rows 0 2>&1 | tail -1 > gpurun_out/bench2_nodp.json; health bench2nodp
for f in gpurun_out/bench2_push.json gpurun_out/bench2_pull.json gpurun_out/bench2_nodp.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d["value"]), d["ms_per_step"], (d.get("e2e") or {}).get("value"), d["clocks"], d["config"]["parallelism"])
except Exception as e:
    print(sys.argv[1], "FAILED", e, open(sys.argv[1]).read()[-600:])
PY
done
timeout 300 $T --master-port 29521 tools/step_profile.py --gpus 2 > gpurun_out/g2_profile_push.md 2>&1; health profile
grep -v "^$" gpurun_out/g2_profile_push.md | tail -30
