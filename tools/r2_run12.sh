set -x
for st in 1 0; do
  VSD_ATTN_STAGGER=$st timeout 300 python tools/gpu_check.py attn_timing 2>&1 | grep "TIME\|EXC" | cut -c1-200
done
timeout 300 python tools/gpu_check.py attn > gpurun_out/r2_attn.txt 2>&1
grep -c PASS gpurun_out/r2_attn.txt; grep "FAIL\|EXC\|DONE" gpurun_out/r2_attn.txt | cut -c1-300 | head -30
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention2 -s 3 -c 1 -f -o gpurun_out/r02_prof_attention2b python tools/ncu_one_attn.py 4096 40 8 > gpurun_out/ncu_attn.log 2>&1
tail -2 gpurun_out/ncu_attn.log
