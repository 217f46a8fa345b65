set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 2>gpurun_out/bench_r1_final.err > gpurun_out/bench_r1_final.json; tail -c 3000 gpurun_out/bench_r1_final.json
python bench.py --steps 20 --warmup 5 --workload cfg3 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_r1_final_cfg3.json
python scripts/measure_small.py 2>&1 | tail -6 > gpurun_out/small_final.txt; cat gpurun_out/small_final.txt
