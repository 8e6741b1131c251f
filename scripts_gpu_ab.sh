#!/bin/bash
# A/B micro-benchmarks + numerics for the round-1 kernel changes (1 GPU). Outputs in gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_interaction_gpu.py tests/test_gemm_gpu.py -x -q 2>&1 | tail -4 > gpurun_out/ab_pytest.log; cat gpurun_out/ab_pytest.log
(echo "## new kernels"; timeout 300 python tools/microbench.py gemm interaction tbe) > gpurun_out/ab_new.md 2>&1
(echo "## legacy interaction, 128x128 GEMM tiles"; TRB_INTERACTION_LEGACY=1 TRB_GEMM_WIDE=0 timeout 300 python tools/microbench.py gemm interaction) > gpurun_out/ab_old.md 2>&1
cat gpurun_out/ab_new.md gpurun_out/ab_old.md
timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/bench1b.json; cut -c1-330 gpurun_out/bench1b.json
for k in tbe_bwd_chunk_kernel tbe_pooled_fwd_chunk_kernel interaction_bwd_pipe_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 3 -c 1 -f -o gpurun_out/ncu_$k python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
