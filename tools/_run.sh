timeout 600 python -m pytest tests/test_gpu_scp.py -x -q --durations=5 2>&1 | tail -40
