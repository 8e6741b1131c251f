#!/bin/bash
mkdir -p gpurun_out
health() { timeout 60 nvidia-smi --query-gpu=index,memory.used --format=csv,noheader || { echo "GPU UNHEALTHY after $1"; exit 7; }; }
timeout 300 python bench.py --steps 20 --warmup 5 --trace-e2e gpurun_out/e2e_trace.json > gpurun_out/bench1_trace.log 2>&1; health bench
grep "^{" gpurun_out/bench1_trace.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench1', round(d['value']), d['ms_per_step'], 'host', round(d['host_enqueue_ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['e2e']['ms_per_step'], 'e2e host', round(d['e2e']['host_enqueue_ms_per_step'],3))"
TRB_BENCH_BIND_NUMA=0 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench1_nobind.log 2>&1; health bench
grep "^{" gpurun_out/bench1_nobind.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench1 nobind', round(d['value']), d['ms_per_step'], 'host', round(d['host_enqueue_ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['e2e']['ms_per_step'], 'e2e host', round(d['e2e']['host_enqueue_ms_per_step'],3))"
head -60 gpurun_out/e2e_trace.json.txt | cut -c1-200
ls -la gpurun_out/e2e_trace.json | cut -c1-100; gzip -f gpurun_out/e2e_trace.json
