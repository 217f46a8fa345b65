set -x
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/measure_strips.py 2>&1 | grep -E "us/step|rror|IPC|n_shapes" 
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 2>gpurun_out/bench_n2.err | tee gpurun_out/bench_n2.json | cut -c1-600
