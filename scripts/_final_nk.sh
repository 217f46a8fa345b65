N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r3k_cfg3_n$N.json 2> gpurun_out/r3k_cfg3_n$N.err; tail -1 gpurun_out/r3k_cfg3_n$N.err | cut -c1-200
timeout 900 $TR bench.py --gpus $N --workload cfg5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r3k_cfg5_n$N.json 2> gpurun_out/r3k_cfg5_n$N.err; tail -1 gpurun_out/r3k_cfg5_n$N.err | cut -c1-200
python - <<P
import json
for W in ("cfg3","cfg5"):
    d=json.loads(open(f'gpurun_out/r3k_{W}_n$N.json').read().strip().splitlines()[-1]); print(W, 'n$N', round(d['ms_per_step'],4), f"{d['value']:.3e}", d.get('strip_counts_ok'), 'e2e', round(d['e2e']['ms_per_step'],3), (d.get('rays') or {}).get('rays_per_sec'))
P
