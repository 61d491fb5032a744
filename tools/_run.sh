mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_r2_reference_arm.json 2>/dev/null; head -c 1200 gpurun_out/bench_r2_reference_arm.json; echo
python bench.py > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['e2e']['value'], d['e2e_tail']['value'], d['cpu_baseline']['value'])
print(d['problems']['car']['kernel_ms'], d['problems']['hopper']['kernel_ms'])
PY
CONFIG5_M=1000000 python examples/config5_gaussian_vs_saa.py > gpurun_out/config5_r2.json 2> gpurun_out/config5.err; tail -2 gpurun_out/config5.err; cat gpurun_out/config5_r2.json
