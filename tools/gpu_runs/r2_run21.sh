mkdir -p gpurun_out
export VSD_WATCHDOG_S=20 VSD_TRACE=1
run() { name=$1; shift; echo "=== $name: $*"; env "$@" timeout 300 python bench.py --config sessions --no-cpu-baseline $EXTRA > gpurun_out/s_$name.json 2> gpurun_out/s_$name.err; echo "rc=$?"; grep -v "^  File\|^    \|Traceback" gpurun_out/s_$name.err | tail -45 | cut -c1-330; head -c 150 gpurun_out/s_$name.json; echo; }
EXTRA="" run default A=1
EXTRA="" run default2 A=1
EXTRA="" run nopdl VSD_PDL=0
EXTRA="--session-lanes 1" run lanes1 A=1
EXTRA="--max-batch 1" run batch1 A=1
