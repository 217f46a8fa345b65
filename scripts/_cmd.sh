set -x
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_fuzz_gpu.py tests/test_strips_gpu.py -m gpu -x -q 2>&1 | tail -3
bash scripts/quick_bench.sh 2>&1 | tail -2
bash scripts/quick_bench.sh --workload cfg3 2>&1 | tail -2
