timeout 400 python tools/kernel_bench.py --md gpurun_out/kernel_roofline_r2.md > gpurun_out/kb.log 2>&1
timeout 700 ncu --set full --clock-control none --import-source on -k regex:'tbe_pooled_fwd|tbe_bwd_chunk|tbe_bwd_span|kjt_route|trb_grad_push|trb_staging|qtbe_fwd|tbe_bwd_build' -c 30 -f -o gpurun_out/ncu_kernels_r2 python tools/kernel_bench.py --iters 1 --warm 0 --no-flush --skip codec,jagged > gpurun_out/ncu_kb.log 2>&1
timeout 120 python -m pytest tests/test_jagged_qcomm_gpu.py -x -q 2>&1 | tail -3
tail -40 gpurun_out/kb.log; tail -3 gpurun_out/ncu_kb.log
