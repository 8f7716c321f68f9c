set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x > gpurun_out/r2_pytest_gpu.txt 2>&1
tail -5 gpurun_out/r2_pytest_gpu.txt
timeout 300 python tools/gpu_check.py s2_sweep > gpurun_out/r2_s2sweep.txt 2>&1
grep -c PASS gpurun_out/r2_s2sweep.txt; grep "FAIL\|EXC\|DONE\|SWEEP" gpurun_out/r2_s2sweep.txt | cut -c1-300 | head -20
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
python -c "
import json;d=json.load(open('gpurun_out/r2_bench_b.json'));print('BENCH',d['value'],d['e2e']['value'],d['e2e']['p50_ms'],d['single_lane']['value'],d['tuning']['table_misses'],d['launches_per_frame'],d['e2e']['output_sha256'],d['roofline']['frac'])"
timeout 900 python bench.py --no-cpu-baseline --paced-frames 0 --steps 60 > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err
python -c "
import json;d=json.load(open('gpurun_out/r2_bench_c.json'));print('BENCH2',d['value'],d['e2e']['value'],d['e2e']['output_sha256'])"
