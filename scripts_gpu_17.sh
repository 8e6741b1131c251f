#!/bin/bash
mkdir -p gpurun_out
health() { timeout 60 nvidia-smi --query-gpu=index,memory.used --format=csv,noheader || { echo "GPU UNHEALTHY after $1"; exit 7; }; }
timeout 400 python -m pytest tests/test_gemm_gpu.py tests/test_interaction_gpu.py tests/test_head_gpu.py -x -q 2>&1 | tail -6; health tests
for b in 0 1; do
TRB_GEMM_RELU_BITS=$b timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench1_bits$b.log 2>&1; health bench$b
grep "^{" gpurun_out/bench1_bits$b.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench1 relu_bits=$b', round(d['value']), d['ms_per_step'], 'host', round(d['host_enqueue_ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['e2e']['ms_per_step'], 'loss', d['loss'])"
done
