# usage: bash scripts/_cmd3.sh <tag> [pytest files...]   — quick parity run + path timing + ncu launch list
TAG=$1; shift
python -m pytest "$@" -x -q -m gpu 2>&1 | tail -6
timeout 300 python scripts/measure_paths.py > gpurun_out/${TAG}_paths.json 2> gpurun_out/${TAG}_paths.err; tail -3 gpurun_out/${TAG}_paths.err
python - <<P
import json
d=json.load(open("gpurun_out/${TAG}_paths.json"))
for k,v in d.items():
    print(k, "identical", v["identical"], "walk", v["walk"]["ms"], "cellpairs", v["cellpairs"]["ms"], v["cellpairs"]["kernels_us"])
P
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/${TAG}_launches.csv python scripts/profile_frame.py --frames 1 > gpurun_out/${TAG}_ncu.log 2>&1
python - <<P
import csv
rows=list(csv.reader(open("gpurun_out/${TAG}_launches.csv")))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
print(" ".join(f"{r[ki].split('(')[0].replace('void ','')}={float(r[vi])/1000:.1f}" for r in rows[hdr+1:][-13:] if len(r)>vi))
P
