#!/bin/bash
mkdir -p gpurun_out
health() { timeout 60 nvidia-smi --query-gpu=index,memory.used --format=csv,noheader || { echo "GPU UNHEALTHY after $1"; exit 7; }; }
timeout 600 python -m pytest tests/test_interaction_gpu.py tests/test_gemm_gpu.py -m gpu -q -x 2>&1 | tail -40 > gpurun_out/g1_pytest.log; tail -5 gpurun_out/g1_pytest.log; health pytest
for g in 1 0; do
  timeout 600 python bench.py --cuda-graphs $g 2>&1 | tail -1 > gpurun_out/bench1e_g$g.json; health bench$g
  python - gpurun_out/bench1e_g$g.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d["value"]), d["ms_per_step"], "host_enqueue_ms", d.get("host_enqueue_ms_per_step"), "launches", d["gpu_launches"], "e2e", round(d["e2e"]["value"]), d["e2e"]["ms_per_step"])
except Exception as e:
    print(sys.argv[1], "FAILED", e, open(sys.argv[1]).read()[-1500:])
PY
done
