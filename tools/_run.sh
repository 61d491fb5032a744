python -m pytest tests/test_gpu_drone.py tests/test_gpu_car.py tests/test_gpu_abi.py -x -q 2>&1 | tail -8
