python -m pytest tests/test_gpu_drone.py tests/test_gpu_car.py tests/test_gpu_abi.py -x -q 2>&1 | tail -4
echo "== own"; SAA_B200_LIB=build/own.so python -m pytest tests/test_gpu_drone.py tests/test_gpu_tail.py -x -q 2>&1 | tail -3
SAA_B200_LIB=build/own.so KB_ONLY=drone python tools/kbench_all.py 2>&1 | grep "drone assemble"
echo "== default"; KB_ONLY=drone python tools/kbench_all.py 2>&1 | grep "drone assemble"
