python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-gather > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -5 gpurun_out/bench_n2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','multi_gpu_parity'):
    print(k, d.get(k))
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
print('kernel_ms', d['roofline']['kernel_ms'], d['per_step_ms'])
PY
