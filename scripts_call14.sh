timeout 300 python -m pytest tests/test_zch_gpu.py -q 2>&1 | grep -v Warning | tail -40
