set -x
mkdir -p gpurun_out
timeout 600 python tools/gpu_check.py gemm attn > gpurun_out/r2_gemm.txt 2>&1
grep -c PASS gpurun_out/r2_gemm.txt; grep "FAIL\|EXC\|DONE" gpurun_out/r2_gemm.txt | cut -c1-300 | head -30
for i in 1 2 3; do
timeout 600 python bench.py --config sessions --no-cpu-baseline > gpurun_out/r2_bench_sessions_$i.json 2> gpurun_out/r2_bench_sessions_$i.err; echo "rc=$?"
tail -2 gpurun_out/r2_bench_sessions_$i.err | cut -c1-200
python -c "
import json;d=json.load(open('gpurun_out/r2_bench_sessions_$i.json'));print('SESS',d['value'],d.get('e2e',{}).get('value'),d.get('tuning'))"
done
timeout 900 python bench.py --no-cpu-baseline --paced-frames 0 > gpurun_out/r2_bench_c512_3.json 2> gpurun_out/r2_bench_c512_3.err
python -c "
import json;d=json.load(open('gpurun_out/r2_bench_c512_3.json'));print('C512',d['value'],d['e2e']['value'],d['e2e']['p50_ms'],d['single_lane']['value'],d['tuning']['table_misses'],d['launches_per_frame'],d['e2e']['output_matches_golden'],d['roofline']['frac'])"
