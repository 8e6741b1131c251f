timeout 300 python tools/kernel_bench.py --only "tbe_bwd phase 2" --skip quant,codec,jagged 2>&1 | grep "tbe_bwd phase 2" | cut -c1-90
timeout 600 python -m pytest tests/test_tbe_gpu.py tests/test_sparse_plane_gpu.py tests/test_zch_gpu.py -q 2>&1 | tail -12
timeout 300 python bench.py --steps 100 --warmup 10 > gpurun_out/r2_b1_c17.json 2> gpurun_out/r2_b1_c17.err; grep '^{' gpurun_out/r2_b1_c17.json | cut -c1-330
