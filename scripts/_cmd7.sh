python bench.py --steps 2 --warmup 3 --no-parity --no-cpu-baseline > gpurun_out/r3f_plain.json 2> gpurun_out/r3f_plain.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3f_launches.csv python bench.py --steps 2 --warmup 3 --no-parity --no-cpu-baseline > gpurun_out/r3f_ncu_launch.log 2>&1
tail -1 gpurun_out/r3f_ncu_launch.log | head -c 300; echo
python scripts/profile_frame.py --frames 1 > gpurun_out/r3f_profile_plain.log 2>&1 &&
timeout 700 ncu --set full --clock-control none --import-source on -s 12 -c 12 -o gpurun_out/r3f_prof -f python scripts/profile_frame.py --frames 1 > gpurun_out/r3f_ncufull.log 2>&1
tail -1 gpurun_out/r3f_ncufull.log
