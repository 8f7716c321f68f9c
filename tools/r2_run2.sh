set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_lanes.py tests/test_gpu_boot.py -x -q > gpurun_out/r2_lanes_tests.txt 2>&1
tail -25 gpurun_out/r2_lanes_tests.txt
python tools/make_tuning_tables.py 512x512x1:1 512x512x1:4 256x256x1:1:cn:kl 2>&1 | tail -5
