timeout 900 python scripts/fuzz_long.py 320 5000 2>&1 | tail -6
