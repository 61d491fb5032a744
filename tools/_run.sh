python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_r2_n8.json 2> gpurun_out/bench_n8.err; tail -5 gpurun_out/bench_n8.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_n8.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','multi_gpu_parity','target_config'):
    print(k, d.get(k))
print('e2e', d['e2e'].get('value'), d['e2e'].get('ms_per_step'))
print('kernel_ms', d['roofline']['kernel_ms'], d['per_step_ms'])
PY
