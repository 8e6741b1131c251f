#!/bin/bash
# 1 GPU: numerics of the new GEMM/colsum paths, micro A/B, bench
mkdir -p gpurun_out
health() { timeout 60 nvidia-smi --query-gpu=index,memory.used --format=csv,noheader || { echo "GPU UNHEALTHY after $1"; exit 7; }; }
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_head_gpu.py -x -q 2>&1 | tail -5; health tests
echo "== colsum v1"; TRB_COLSUM=1 timeout 120 python tools/microbench.py colsum 2>&1 | tail -10
echo "== colsum v2"; timeout 120 python tools/microbench.py colsum 2>&1 | tail -10; health colsum
echo "== gemm split wide=0"; TRB_GEMM_SPLIT_WIDE=0 timeout 200 python tools/microbench.py gemm 2>&1 | grep wgrad
echo "== gemm split wide=1"; timeout 200 python tools/microbench.py gemm 2>&1 | tee gpurun_out/microbench_gemm.md | grep wgrad; health gemm
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench1_c30.log 2>&1; health bench
grep "^{" gpurun_out/bench1_c30.log | tail -1 > gpurun_out/bench1_c30.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench1_c30.json")); print("bench1", round(d["value"]), d["ms_per_step"], "host", d.get("host_enqueue_ms_per_step"), "launches", d["gpu_launches"], "e2e", round(d["e2e"]["value"]))
PY
