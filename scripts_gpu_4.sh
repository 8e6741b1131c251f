#!/bin/bash
# 1 GPU: colsum v3, e2e host profile, then (risky, last, under timeout) the CTA-pair GEMM
mkdir -p gpurun_out
health() { timeout 60 nvidia-smi --query-gpu=index,memory.used --format=csv,noheader || { echo "GPU UNHEALTHY after $1"; exit 7; }; }
timeout 300 python -m pytest tests/test_gemm_gpu.py -x -q -k "colsum or linear or split" 2>&1 | tail -3; health tests
echo "== colsum v3"; timeout 120 python tools/microbench.py colsum 2>&1 | tail -9; health colsum
timeout 300 python bench.py --steps 30 --warmup 5 --profile-host > gpurun_out/bench1_c31.log 2>&1; health bench
grep "^{" gpurun_out/bench1_c31.log | tail -1 > gpurun_out/bench1_c31.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench1_c31.json")); print("bench1", round(d["value"]), d["ms_per_step"], "host", d.get("host_enqueue_ms_per_step"), "launches", d["gpu_launches"], "e2e", round(d["e2e"]["value"]))
PY
grep -A 48 "host profile: TrainPipelineSparseDist.progress (10 steps), sorted by cumulative" gpurun_out/bench1_c31.log | cut -c1-180
echo "== pair kernel tests"
timeout 240 python -m pytest tests/test_gemm_gpu.py -x -q -k "cta_pair" 2>&1 | tail -15; health pair
echo "== gemm pair=1"; TRB_GEMM_PAIR=1 timeout 200 python tools/microbench.py gemm 2>&1 | tee gpurun_out/microbench_gemm_pair.md | tail -22; health gemm
TRB_GEMM_PAIR=1 timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench1_pair.log 2>&1; health bench
grep "^{" gpurun_out/bench1_pair.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench1 pair', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']))"
