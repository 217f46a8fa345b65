set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash scripts/quick_bench.sh 2>&1 | tail -3
bash scripts/quick_bench.sh --workload cfg3 2>&1 | tail -3
