#!/bin/bash
mkdir -p gpurun_out
health() { timeout 60 nvidia-smi --query-gpu=index,memory.used --format=csv,noheader || { echo "GPU UNHEALTHY after $1"; exit 7; }; }
timeout 300 python -m pytest tests/test_tbe_gpu.py tests/test_gemm_gpu.py -q -k "stochastic or crossnet" 2>&1 | tail -25; health tests
