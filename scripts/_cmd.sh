set -x
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python scripts/quick_bench_fmt.py
python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python scripts/quick_bench_fmt.py
for n in 4; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 2>gpurun_out/bench_n$n.err > gpurun_out/bench_n$n.json; cut -c1-230 gpurun_out/bench_n$n.json
done
