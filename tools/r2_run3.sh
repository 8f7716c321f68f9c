set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/r2_pytest_gpu.txt 2>&1
tail -15 gpurun_out/r2_pytest_gpu.txt
