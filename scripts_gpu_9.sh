#!/bin/bash
mkdir -p gpurun_out
health() { timeout 60 nvidia-smi --query-gpu=index,memory.used --format=csv,noheader || { echo "GPU UNHEALTHY after $1"; exit 7; }; }
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_head_gpu.py -x -q 2>&1 | tail -4; health tests
echo "== gemmx"; timeout 300 python tools/microbench.py gemmx 2>&1 | tee gpurun_out/microbench_gemmx3.md | tail -16; health gemmx
echo "== gemm"; timeout 200 python tools/microbench.py gemm 2>&1 | tee gpurun_out/microbench_gemm_p2.md | tail -20; health gemm
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench1_c35.log 2>&1; health bench
grep "^{" gpurun_out/bench1_c35.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench1', round(d['value']), d['ms_per_step'], 'host', round(d['host_enqueue_ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['e2e']['ms_per_step'])"
