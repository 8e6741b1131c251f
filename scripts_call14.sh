timeout 300 python tools/kernel_bench.py --skip sparse,codec,jagged 2>&1 | grep "qtbe" | cut -c1-100
timeout 600 python -m pytest tests/test_quant_gpu.py tests/test_quant_sharded_gpu.py tests/test_gemm_gpu.py -q 2>&1 | tail -3
