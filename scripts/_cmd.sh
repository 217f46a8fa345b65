timeout 300 python -m pytest tests/test_lines_gpu.py -m gpu -x -q -k platformer 2>&1 | tail -40
