TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 8 --workload cfg5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r3k_cfg5_n8.json 2> gpurun_out/r3k_cfg5_n8.err; tail -2 gpurun_out/r3k_cfg5_n8.err | cut -c1-300
python - <<'P'
import json
d=json.loads(open('gpurun_out/r3k_cfg5_n8.json').read().strip().splitlines()[-1]); print('cfg5 n8', d['ms_per_step'], d['value'], d.get('strip_counts_ok'), d.get('kernels_us_rank0'), d['e2e'].get('ms_per_step'), d.get('rays'))
P
