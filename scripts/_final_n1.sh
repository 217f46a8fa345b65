# end-of-round single-GPU batch: the whole GPU suite, then every config's bench line
python -m pytest tests -q -m gpu 2>&1 | tail -4
for W in cfg3 cfg2 cfg1 cfg4; do
  timeout 900 python bench.py --workload $W > gpurun_out/r3k_${W}_n1.json 2> gpurun_out/r3k_${W}_n1.err; tail -2 gpurun_out/r3k_${W}_n1.err
done
timeout 1500 python bench.py --workload cfg5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r3k_cfg5_n1.json 2> gpurun_out/r3k_cfg5_n1.err; tail -2 gpurun_out/r3k_cfg5_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r3k_reference.json 2> gpurun_out/r3k_reference.err; tail -2 gpurun_out/r3k_reference.err
python - <<'P'
import json
for W in ("cfg3","cfg2","cfg1","cfg4","cfg5"):
    try:
        d=json.loads(open(f"gpurun_out/r3k_{W}_n1.json").read().strip().splitlines()[-1])
        print(W, round(d["ms_per_step"],4), f'{d["value"]:.3e}', "parity", d.get("parity_counts_ok"), "e2e", round(d["e2e"]["ms_per_step"],3), d.get("kernels_us"), (d.get("rays") or {}).get("rays_per_sec"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(W, "ERR", e)
try:
    d=json.loads(open("gpurun_out/r3k_reference.json").read().strip().splitlines()[-1]); print("reference", d["value"], d["ms_per_step"], d["cpu_baseline"]["cores"])
except Exception as e: print("reference ERR", e)
P
