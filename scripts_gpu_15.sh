#!/bin/bash
mkdir -p gpurun_out
health() { timeout 60 nvidia-smi --query-gpu=index,memory.used --format=csv,noheader || { echo "GPU UNHEALTHY after $1"; exit 7; }; }
timeout 600 python -m pytest tests/test_tbe_gpu.py tests/test_uvm_cache.py -x -q -m gpu 2>&1 | tail -5; health tests
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench1_c45.log 2>&1; health bench
grep "^{" gpurun_out/bench1_c45.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench1', round(d['value']), d['ms_per_step'], 'host', round(d['host_enqueue_ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['e2e']['ms_per_step'], 'e2e host', round(d['e2e']['host_enqueue_ms_per_step'],3), d['config'].get('cpu_binding'))"
