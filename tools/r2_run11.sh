set -x
mkdir -p gpurun_out
for poly in 0 4 3; do
  VSD_ATTN_POLY=$poly timeout 300 python tools/gpu_check.py attn_timing 2>&1 | grep "TIME\|EXC" | cut -c1-200
done
timeout 300 python tools/gpu_check.py attn > gpurun_out/r2_attn.txt 2>&1
grep -c PASS gpurun_out/r2_attn.txt; grep "FAIL\|EXC\|DONE" gpurun_out/r2_attn.txt | cut -c1-300 | head -30
