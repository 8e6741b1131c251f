timeout 600 python -m pytest tests/test_gemm_gpu.py -q -x 2>&1 | tail -12
timeout 400 python bench.py --steps 100 --warmup 10 --phase-times 2>&1 | grep "phase_ms\|^{" | cut -c1-330
TRB_EPI_COLSUM=0 timeout 400 python bench.py --steps 100 --warmup 10 --no-e2e --phase-times 2>&1 | grep "phase_ms\|^{" | cut -c1-230
