set -x
mkdir -p gpurun_out
python tools/gpu_check.py gemm > gpurun_out/r2_gemm.txt 2>&1
grep -c PASS gpurun_out/r2_gemm.txt; grep "FAIL\|EXC\|DONE" gpurun_out/r2_gemm.txt | head -30
export VSD_TUNING_TABLES=0
python tools/gpu_pipeline_check.py 512x512x1 2>&1 | tail -12
python tools/profile_frame.py --frames 3 --save-tuning > gpurun_out/r2_tune3.log 2>&1
python tools/profile_frame.py --frames 3 --sections 4 > gpurun_out/r2_sections4_lnfuse.txt 2>&1
grep "SECTIONS\|launches" gpurun_out/r2_sections4_lnfuse.txt
grep "splits=[2-9]" gpurun_out/tune_cache_512x512x1.txt | awk '{print $6}' | sort | uniq -c
