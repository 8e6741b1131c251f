#!/bin/bash
# 1-GPU final pass: smoke, full GPU test suite, bench (default flags)
mkdir -p gpurun_out
health() { timeout 60 nvidia-smi --query-gpu=index,memory.used --format=csv,noheader || { echo "GPU UNHEALTHY after $1"; exit 7; }; }
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log; health smoke
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6; health tests
timeout 300 python bench.py > gpurun_out/bench1_final.log 2>&1; health bench
grep "^{" gpurun_out/bench1_final.log | tail -1 > gpurun_out/bench1_final.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench1_final.json")); print("bench1", round(d["value"]), d["ms_per_step"], "host", d.get("host_enqueue_ms_per_step"), "launches", d["gpu_launches"], "e2e", round(d["e2e"]["value"]), d["e2e"]["ms_per_step"], d["clocks"])
PY
