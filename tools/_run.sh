echo "== tile32 tests"; SAA_DRONE_TILE=32 python -m pytest tests/test_gpu_drone.py tests/test_gpu_tail.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3
echo "== tile32"; SAA_DRONE_TILE=32 KB_ONLY=drone python tools/kbench_all.py 2>&1 | grep "drone assemble"
echo "== tile16"; KB_ONLY=drone python tools/kbench_all.py 2>&1 | grep "drone assemble"
