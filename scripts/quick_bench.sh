#!/bin/bash
# quick device-resident timing of the bench workload (no CPU baseline)
python bench.py --steps 20 --warmup 5 --no-cpu-baseline ${@} 2>/dev/null | python scripts/quick_bench_fmt.py
