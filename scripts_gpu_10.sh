#!/bin/bash
mkdir -p gpurun_out
health() { timeout 60 nvidia-smi --query-gpu=index,memory.used --format=csv,noheader || { echo "GPU UNHEALTHY after $1"; exit 7; }; }
timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -3; health tests
echo "== gemmx"; timeout 300 python tools/microbench.py gemmx 2>&1 | grep "mask\|MN-major B, plain" ; health gemmx
for g in -1; do
timeout 300 python bench.py --steps 30 --warmup 5 --cuda-graphs $g > gpurun_out/bench1_g$g.log 2>&1; health bench$g
grep "^{" gpurun_out/bench1_g$g.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench1 graphs=$g', round(d['value']), d['ms_per_step'], 'host', round(d['host_enqueue_ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['e2e']['ms_per_step'])"
done
true
true
