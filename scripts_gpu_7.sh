#!/bin/bash
mkdir -p gpurun_out
health() { timeout 60 nvidia-smi --query-gpu=index,memory.used --format=csv,noheader || { echo "GPU UNHEALTHY after $1"; exit 7; }; }
timeout 600 python -m pytest tests/test_tbe_gpu.py tests/test_gemm_gpu.py -x -q 2>&1 | tail -4; health tests
echo "== colsum v3b"; timeout 120 python tools/microbench.py colsum 2>&1 | tail -9; health colsum
for pf in 0 1; do
TRB_BWD_PREFETCH=$pf timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e > gpurun_out/bench1_pf$pf.log 2>&1; health bench$pf
grep "^{" gpurun_out/bench1_pf$pf.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench1 prefetch=$pf', round(d['value']), d['ms_per_step'], 'host', round(d['host_enqueue_ms_per_step'],3))"
done
echo "== tbe microbench"; TRB_BWD_PREFETCH=0 timeout 200 python tools/microbench.py tbe 2>&1 | grep bwd; timeout 200 python tools/microbench.py tbe 2>&1 | grep bwd; health tbe
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05_pair --launch-skip 2 -c 1 -f -o gpurun_out/ncu_gemm_pair python tools/gemm_one.py 32768 1024 1024 512 1 > gpurun_out/ncu_gemm_pair.log 2>&1; tail -2 gpurun_out/ncu_gemm_pair.log; health ncu
ls -la gpurun_out/*.ncu-rep
