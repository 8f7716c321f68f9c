export VSD_TMA_S2=1
for i in 1 2 3; do
timeout 300 python tools/gpu_pipeline_check.py 360x640x1 > gpurun_out/r2_s2_360_$i.txt 2>&1; echo "rc=$?"; grep "teacher\|free-running\|PSNR\|TIMING\|rror\|nan" gpurun_out/r2_s2_360_$i.txt | cut -c1-200 | head -12
done
timeout 300 python tools/gpu_pipeline_check.py 512x512x1 > gpurun_out/r2_s2_512.txt 2>&1; echo "rc=$?"; grep "PSNR\|TIMING\|rror\|nan" gpurun_out/r2_s2_512.txt | cut -c1-200 | head
unset VSD_TMA_S2
timeout 300 python tools/gpu_pipeline_check.py 360x640x1 > gpurun_out/r2_s2off_360.txt 2>&1; echo "rc=$?"; grep "PSNR\|TIMING\|rror\|nan" gpurun_out/r2_s2off_360.txt | cut -c1-200 | head
