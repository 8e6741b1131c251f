#!/bin/bash
mkdir -p gpurun_out
health() { timeout 60 nvidia-smi --query-gpu=index,memory.used --format=csv,noheader || { echo "GPU UNHEALTHY after $1"; exit 7; }; }
TRB_COLSUM=4 timeout 300 python -m pytest tests/test_gemm_gpu.py -x -q -k "colsum or linear" 2>&1 | tail -3; health tests
echo "== colsum v3"; timeout 120 python tools/microbench.py colsum 2>&1 | tail -9
echo "== colsum v4"; TRB_COLSUM=4 timeout 120 python tools/microbench.py colsum 2>&1 | tail -9; health colsum
nvidia-smi topo -m 2>/dev/null | head -6; lscpu | grep -i "numa\|model name\|^CPU(s)" | head -8
for v in 3 4 4; do
TRB_COLSUM=$v timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench1_cs$v.log 2>&1; health bench$v
grep "^{" gpurun_out/bench1_cs$v.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench1 colsum=$v', round(d['value']), d['ms_per_step'], 'host', round(d['host_enqueue_ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['e2e']['ms_per_step'], 'e2e host', round(d['e2e']['host_enqueue_ms_per_step'],3))"
done
