set -x
export RSV_NO_GRAPH=1
python scripts/profile_frame.py --frames 4 > gpurun_out/r1_full_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on --launch-skip 33 --launch-count 11 -o gpurun_out/r1_frame_full -f python scripts/profile_frame.py --frames 4 > gpurun_out/r1_full_ncu.log 2>&1
tail -2 gpurun_out/r1_full_ncu.log
