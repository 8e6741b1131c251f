#!/bin/bash
# 1-GPU validation + profiling pass (run under gpurun). Outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" 
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/bench1.json; cut -c1-600 gpurun_out/bench1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 300 -c 220 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/ncu_launch.log 2>&1
for k in gemm_bf16_tcgen05_kernel tbe_bwd_chunk_kernel interaction_bwd_kernel tbe_pooled_fwd_chunk_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 12 -c 1 -f -o gpurun_out/ncu_$k python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out | tail -12
