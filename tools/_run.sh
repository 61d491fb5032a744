python -m pytest tests/test_gpu_car.py -x -q 2>&1 | tail -3
echo "== default (copy templated, gform 1)"; KB_ONLY=car python tools/kbench_all.py 2>&1 | grep "car assemble"
for v in c00 c10 c11 c01B; do echo "== $v"; SAA_B200_LIB=build/$v.so KB_ONLY=car python tools/kbench_all.py 2>&1 | grep "car assemble"; done
