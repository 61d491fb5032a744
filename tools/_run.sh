echo "== car32 tests"; SAA_CAR_TILE=32 python -m pytest tests/test_gpu_car.py tests/test_gpu_tail.py -x -q 2>&1 | tail -3
echo "== car32"; SAA_CAR_TILE=32 KB_ONLY=car python tools/kbench_all.py 2>&1 | grep "car assemble"
echo "== car16"; KB_ONLY=car python tools/kbench_all.py 2>&1 | grep "car assemble"
