timeout 400 python tools/kernel_bench.py --md gpurun_out/kernel_roofline_r2.md > gpurun_out/kb.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tbe_pooled_fwd|tbe_bwd_chunk|tbe_bwd_span|kjt_route|trb_grad_push|trb_staging|qtbe_fwd|tbe_bwd_build' -c 30 -f -o gpurun_out/ncu_kernels_r2 python tools/kernel_bench.py --iters 1 --warm 0 --no-flush --skip codec,jagged > gpurun_out/ncu_kb.log 2>&1
timeout 300 python bench.py --steps 100 --warmup 10 > gpurun_out/r2_b1_c14.json 2> gpurun_out/r2_b1_c14.err
TRB_SPARSE_GRAPHS=0 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 400 -c 300 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 3 --warmup 3 --no-e2e --cuda-graphs 0 > gpurun_out/l.log 2>&1
tail -40 gpurun_out/kb.log; tail -3 gpurun_out/ncu_kb.log; cat gpurun_out/r2_b1_c14.json | grep '^{' | cut -c1-400
