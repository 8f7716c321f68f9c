export VSD_TUNING_TABLES=0
timeout 200 python tools/gpu_pipeline_check.py 512x512x1 > gpurun_out/r2_pipe_a.txt 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2_pipe_a.txt | cut -c1-300
VSD_ATTN_V2=0 timeout 200 python tools/gpu_pipeline_check.py 512x512x1 > gpurun_out/r2_pipe_b.txt 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2_pipe_b.txt | cut -c1-300
