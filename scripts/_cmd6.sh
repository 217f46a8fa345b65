# N=2: strip validation on two real GPUs + bench lines (config 3 weak scaling, config 5 at 4 M shapes)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR scripts/validate_strips_multi.py --shapes 500000 --frames 3 > gpurun_out/r3e_validate_n2.json 2> gpurun_out/r3e_validate_n2.err; tail -2 gpurun_out/r3e_validate_n2.err; tail -c 600 gpurun_out/r3e_validate_n2.json
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r3e_cfg3_n2.json 2> gpurun_out/r3e_cfg3_n2.err; tail -2 gpurun_out/r3e_cfg3_n2.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r3e_cfg3_n2.json').read().strip().splitlines()[-1]); print('cfg3 n2', d['ms_per_step'], d['value'], d.get('strip_counts_ok'), d.get('kernels_us_rank0'), d['e2e'])
P
