python -m pytest tests -q -m gpu 2>&1 | tail -3
for W in cfg3 cfg2; do
  timeout 900 python bench.py --workload $W > gpurun_out/r3n_${W}_n1.json 2> gpurun_out/r3n_${W}_n1.err; tail -1 gpurun_out/r3n_${W}_n1.err
done
python - <<'P'
import json
for W in ("cfg3","cfg2"):
    d=json.loads(open(f"gpurun_out/r3n_{W}_n1.json").read().strip().splitlines()[-1])
    print(W, round(d["ms_per_step"],4), f'{d["value"]:.3e}', "parity", d.get("parity_counts_ok"), "e2e", round(d["e2e"]["ms_per_step"],3), d.get("kernels_us"), d["roofline"]["kernel"], round(d["roofline"]["frac"],4), d["pair_tests_per_sec"], d["candidate_path"]["literal_walk_ms_per_step"])
P
python bench.py --steps 2 --warmup 3 --no-parity --no-cpu-baseline > gpurun_out/r3n_plain.json 2> gpurun_out/r3n_plain.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3n_launches.csv python bench.py --steps 2 --warmup 3 --no-parity --no-cpu-baseline > gpurun_out/r3n_ncu_launch.log 2>&1
