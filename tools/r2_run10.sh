set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention2 -s 3 -c 1 -f -o gpurun_out/r02_prof_attention2 python tools/ncu_one_attn.py 4096 40 8 > gpurun_out/ncu_attn.log 2>&1
tail -2 gpurun_out/ncu_attn.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 2 -c 1 -f -o gpurun_out/r02_prof_gemm_4096x320x2880 python tools/ncu_one_gemm.py 4096 320 320 9 96 1 769 1 > gpurun_out/ncu_g1.log 2>&1
tail -2 gpurun_out/ncu_g1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 2 -c 1 -f -o gpurun_out/r02_prof_gemm_64x1280x11520 python tools/ncu_one_gemm.py 64 1280 1280 9 96 8 4097 2 > gpurun_out/ncu_g2.log 2>&1
tail -2 gpurun_out/ncu_g2.log
timeout 300 python tools/gpu_check.py attn > gpurun_out/r2_attn.txt 2>&1
grep -c PASS gpurun_out/r2_attn.txt; grep "FAIL\|EXC\|DONE" gpurun_out/r2_attn.txt | cut -c1-300 | head -30
timeout 600 python tools/gpu_check.py tuner_sweep > gpurun_out/r2_sweep.txt 2>&1
grep -c PASS gpurun_out/r2_sweep.txt; grep "FAIL\|EXC\|DONE\|SWEEP" gpurun_out/r2_sweep.txt | cut -c1-300 | head -30
timeout 600 python tools/gpu_check.py gemm > gpurun_out/r2_gemm.txt 2>&1
grep -c PASS gpurun_out/r2_gemm.txt; grep "FAIL\|EXC\|DONE" gpurun_out/r2_gemm.txt | cut -c1-300 | head -30
