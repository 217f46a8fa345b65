run() { # tag, env...
  TAG=$1; shift
  env "$@" timeout 300 python scripts/measure_frame.py > gpurun_out/r3c_$TAG.json 2> gpurun_out/r3c_$TAG.err; echo "$TAG $(cat gpurun_out/r3c_$TAG.json)"; tail -2 gpurun_out/r3c_$TAG.err
}
run base A=1
run scan16 RSV_LIB_PATH=resolv_b200/libresolv_b200_scan16.so
run n33 RSV_NARROW_BLOCKS0=3 RSV_NARROW_BLOCKS1=3
run n22 RSV_NARROW_BLOCKS0=2 RSV_NARROW_BLOCKS1=2
run n32 RSV_NARROW_BLOCKS0=3 RSV_NARROW_BLOCKS1=2
run n42 RSV_NARROW_BLOCKS0=4 RSV_NARROW_BLOCKS1=2
run n52 RSV_NARROW_BLOCKS0=5 RSV_NARROW_BLOCKS1=2
run n23 RSV_NARROW_BLOCKS0=2 RSV_NARROW_BLOCKS1=3
run n88 RSV_NARROW_BLOCKS0=8 RSV_NARROW_BLOCKS1=8
