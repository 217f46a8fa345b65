TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r3n_cfg3_n8.json 2> gpurun_out/r3n_cfg3_n8.err; tail -1 gpurun_out/r3n_cfg3_n8.err | cut -c1-200
python - <<'P'
import json
d=json.loads(open('gpurun_out/r3n_cfg3_n8.json').read().strip().splitlines()[-1]); print('cfg3 n8', d['ms_per_step'], d['value'], d.get('strip_counts_ok'), d.get('kernels_us_rank0'), d['e2e']['ms_per_step'])
P
