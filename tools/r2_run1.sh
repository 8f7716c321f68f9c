set -x
mkdir -p gpurun_out
python tools/gpu_check.py norm > gpurun_out/r2_norm.txt 2>&1
tail -5 gpurun_out/r2_norm.txt
python tools/profile_frame.py --frames 3 --save-tuning > gpurun_out/r2_tune.log 2>&1
python tools/profile_frame.py --frames 3 --sections 4 > gpurun_out/r2_sections4_gn_cluster.txt 2>&1
VSD_GN_MODE=1 python tools/profile_frame.py --frames 3 --sections 4 > gpurun_out/r2_sections4_gn_legacy.txt 2>&1
grep SECTIONS gpurun_out/r2_sections4_gn_cluster.txt gpurun_out/r2_sections4_gn_legacy.txt
python tools/gpu_pipeline_check.py 512x512x1 2>&1 | tail -12
