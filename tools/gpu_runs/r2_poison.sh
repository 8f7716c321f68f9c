mkdir -p gpurun_out
export VSD_POISON=1 VSD_WATCHDOG_S=60
timeout 900 python tools/gpu_check.py gemm tuner_sweep s2_sweep > gpurun_out/r02_poison_sweeps.txt 2>&1
echo "rc=$?"; grep -c PASS gpurun_out/r02_poison_sweeps.txt; grep "FAIL\|EXC\|DONE" gpurun_out/r02_poison_sweeps.txt | cut -c1-400 | head -40
