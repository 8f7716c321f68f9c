set -x
timeout 600 python tools/gpu_check.py gemm > gpurun_out/r2_gemm.txt 2>&1
grep -c PASS gpurun_out/r2_gemm.txt; grep "FAIL\|EXC\|DONE" gpurun_out/r2_gemm.txt | cut -c1-300 | head -30
export VSD_TUNING_TABLES=0
timeout 600 python tools/gpu_pipeline_check.py 512x512x1 2>&1 | grep "TIMING\|launches\|PSNR\|free-running"
timeout 900 python bench.py --lanes 6 --steps 60 --no-cpu-baseline --paced-frames 0 > gpurun_out/r2_bench_l6.json 2> gpurun_out/r2_bench_l6.err
python -c "
import json;d=json.load(open('gpurun_out/r2_bench_l6.json'));print('LANES',6,d['value'],d['e2e']['value'],d['e2e']['p50_ms'],d['single_lane']['value'],d['tuning']['table_misses'],d['launches_per_frame'])"
