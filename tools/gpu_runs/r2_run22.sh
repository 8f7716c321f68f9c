mkdir -p gpurun_out
export VSD_WATCHDOG_S=20 VSD_TRACE=1
run() { name=$1; shift; echo "=== $name: $*"; env "$@" timeout 300 python bench.py --config sessions --no-cpu-baseline $EXTRA > gpurun_out/s_$name.json 2> gpurun_out/s_$name.err; echo "rc=$?"; grep -v "^  File\|^    \|Traceback" gpurun_out/s_$name.err | tail -45 | cut -c1-330; head -c 150 gpurun_out/s_$name.json; echo; }
EXTRA="" run d1 A=1
EXTRA="" run d2 A=1
EXTRA="" run d3 A=1
EXTRA="--max-batch 1" run batch1 A=1
unset VSD_TRACE
EXTRA="" run d4 A=1
EXTRA="" run d5 A=1
timeout 600 python tools/gpu_check.py gemm > gpurun_out/r2_gemm.txt 2>&1
echo "gemm PASS count: $(grep -c PASS gpurun_out/r2_gemm.txt)"; grep "FAIL\|EXC\|DONE" gpurun_out/r2_gemm.txt | cut -c1-300 | head -30
