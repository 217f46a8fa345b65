set -x
timeout 600 python -m pytest tests/test_lines_gpu.py tests/test_host_gpu.py -m gpu -x -q 2>&1 | tail -3
python scripts/measure_rays.py 2>&1 | tail -3
