#!/bin/bash
mkdir -p gpurun_out
health() { timeout 60 nvidia-smi --query-gpu=index,memory.used --format=csv,noheader || { echo "GPU UNHEALTHY after $1"; exit 7; }; }
timeout 300 python -m pytest tests/test_gemm_gpu.py -x -q -k "colsum or linear" 2>&1 | tail -3; health tests
echo "== colsum v1"; TRB_COLSUM=1 timeout 120 python tools/microbench.py colsum 2>&1 | tail -10
echo "== colsum v3"; timeout 120 python tools/microbench.py colsum 2>&1 | tail -10; health colsum
echo "== gemmx"; timeout 300 python tools/microbench.py gemmx 2>&1 | tee gpurun_out/microbench_gemmx.md | tail -18; health gemmx
