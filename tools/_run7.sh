timeout 600 python -m pytest tests/test_gpu_dp.py -q -m gpu 2>&1 | grep -E "AssertionError: |passed|failed|^E  " | cut -c1-600 | head -20
