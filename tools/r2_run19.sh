export VSD_WATCHDOG_S=25
run() { echo "=== $*"; env "$@" timeout 400 python bench.py --config sessions --no-cpu-baseline > gpurun_out/s.json 2> gpurun_out/s.err; echo "rc=$?"; tail -2 gpurun_out/s.err | cut -c1-250; head -c 200 gpurun_out/s.json; echo; }
run A=1
run VSD_ATTN_V2=0
run VSD_LN_FUSE=0
run VSD_FF_OUT_FUSE=0 VSD_TMA_S2=0
run VSD_LN_FUSE=0 VSD_FF_OUT_FUSE=0 VSD_TMA_S2=0 VSD_ATTN_V2=0
