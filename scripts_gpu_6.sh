#!/bin/bash
mkdir -p gpurun_out
health() { timeout 60 nvidia-smi --query-gpu=index,memory.used --format=csv,noheader || { echo "GPU UNHEALTHY after $1"; exit 7; }; }
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_head_gpu.py tests/test_interaction_gpu.py -x -q 2>&1 | tail -12; health tests
echo "== gemmx"; timeout 300 python tools/microbench.py gemmx 2>&1 | tee gpurun_out/microbench_gemmx2.md | tail -18; health gemmx
echo "== gemm pair=0"; timeout 200 python tools/microbench.py gemm 2>&1 | tee gpurun_out/microbench_gemm_p0.md | tail -20; health gemm0
echo "== gemm pair=1"; TRB_GEMM_PAIR=1 timeout 200 python tools/microbench.py gemm 2>&1 | tee gpurun_out/microbench_gemm_p1.md | tail -20; health gemm1
for pr in 0 1; do
TRB_GEMM_PAIR=$pr timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench1_p$pr.log 2>&1; health bench$pr
grep "^{" gpurun_out/bench1_p$pr.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench1 pair=$pr', round(d['value']), d['ms_per_step'], 'host', round(d['host_enqueue_ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['e2e']['ms_per_step'])"
done
