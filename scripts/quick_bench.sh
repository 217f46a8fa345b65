#!/bin/bash
# quick device-resident timing of the bench workload (no CPU baseline)
python bench.py --steps 20 --warmup 5 --no-cpu-baseline ${@} 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), d['kernels_us'], round(d['e2e']['ms_per_step'],3))"
