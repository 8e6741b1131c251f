timeout 1700 python -m pytest tests/ -q -m gpu -x 2>&1 | tail -15
