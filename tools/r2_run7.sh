set -x
mkdir -p gpurun_out
for poly in 4 0 3 2; do
  VSD_ATTN_POLY=$poly timeout 300 python tools/gpu_check.py attn > gpurun_out/r2_attn_poly$poly.txt 2>&1
  grep -c PASS gpurun_out/r2_attn_poly$poly.txt; grep "FAIL\|EXC\|DONE\|TIME" gpurun_out/r2_attn_poly$poly.txt | head -30
done
VSD_ATTN_V2=0 timeout 300 python tools/gpu_check.py attn > gpurun_out/r2_attn_v1.txt 2>&1
grep "FAIL\|EXC\|DONE\|TIME" gpurun_out/r2_attn_v1.txt | head
timeout 300 python tools/gpu_check.py gemm > gpurun_out/r2_gemm.txt 2>&1
grep -c PASS gpurun_out/r2_gemm.txt; grep "FAIL\|EXC\|DONE" gpurun_out/r2_gemm.txt | head -30
export VSD_TUNING_TABLES=0
VSD_LN_FUSE=0 timeout 600 python tools/gpu_pipeline_check.py 512x512x1 2>&1 | grep "TIMING\|launches\|PSNR"
timeout 600 python tools/gpu_pipeline_check.py 512x512x1 2>&1 | grep "TIMING\|launches\|PSNR\|free-running"
