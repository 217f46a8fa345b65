bash scripts/_cmd.sh 2
python scripts/measure_pcie.py > gpurun_out/r2o_pcie_n1.json 2> gpurun_out/r2o_pcie_n1.err; tail -1 gpurun_out/r2o_pcie_n1.json
timeout 1200 python bench.py --workload cfg5 --steps 10 --warmup 3 --no-parity --no-cpu-baseline > gpurun_out/r2o_cfg5_n1.json 2> gpurun_out/r2o_cfg5_n1.err; tail -3 gpurun_out/r2o_cfg5_n1.err
timeout 900 python bench.py --workload cfg5 --shapes 2000000 --rays 1000000 --steps 10 --warmup 3 > gpurun_out/r2o_cfg5_2M_n1.json 2> gpurun_out/r2o_cfg5_2M_n1.err; tail -3 gpurun_out/r2o_cfg5_2M_n1.err
python - <<'P'
import json
for f in ("r2o_cfg5_n1","r2o_cfg5_2M_n1"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],4), d.get("kernels_us"), "e2e", round(d["e2e"]["ms_per_step"],3), "ok", d.get("parity_counts_ok"), d.get("rays"))
    except Exception as e: print(f, "ERR", e)
P
