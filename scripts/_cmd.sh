RSV_NO_GRAPH=1 python scripts/profile_frame.py --frames 3 2>&1 | tail -4
RSV_NO_GRAPH=1 python scripts/profile_frame.py --frames 3 --workload cfg3 2>&1 | tail -4
