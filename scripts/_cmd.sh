set -x
python -m pytest tests -m gpu -x -q -k "not fuzz" 2>&1 | tail -5
python scripts/measure_small.py 2>&1 | tail -8
RSV_NO_GRAPH=1 python scripts/measure_small.py 2>&1 | tail -8
bash scripts/quick_bench.sh 2>&1 | tail -3
