for RB in 1 2 3 4 8; do
RSV_RANK_BLOCKS=$RB timeout 300 python scripts/measure_paths.py > gpurun_out/r2y_paths_$RB.json 2> gpurun_out/r2y_paths_$RB.err
python - <<P
import json
d=json.load(open("gpurun_out/r2y_paths_$RB.json"))
for k,v in d.items():
    print("RB=$RB", k, "identical", v["identical"], "cellpairs", v["cellpairs"]["ms"], v["cellpairs"]["kernels_us"])
P
done
