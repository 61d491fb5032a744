python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json | head -c 6000
