mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x > gpurun_out/r2_pytest_gpu.txt 2>&1
tail -5 gpurun_out/r2_pytest_gpu.txt
for i in 1 2; do
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_c512_$i.json 2> gpurun_out/r2_bench_c512_$i.err
tail -2 gpurun_out/r2_bench_c512_$i.err | cut -c1-300
python -c "
import json;d=json.load(open('gpurun_out/r2_bench_c512_$i.json'));print('C512',d['value'],d['e2e']['value'],d['e2e']['p50_ms'],d['single_lane']['value'],d['tuning']['table_misses'],d['launches_per_frame'],d['e2e']['output_sha256'],d['e2e']['output_matches_golden'],d['roofline']['frac'])"
done
timeout 900 python bench.py --config c768b4 --no-cpu-baseline > gpurun_out/r2_bench_c768.json 2> gpurun_out/r2_bench_c768.err
python -c "
import json;d=json.load(open('gpurun_out/r2_bench_c768.json'));print('C768',d['value'],d['e2e']['value'],d['tuning']['table_misses'],d['launches_per_frame'],d['e2e']['output_sha256'],d['e2e'].get('output_matches_golden'),d['roofline']['frac'])"
for i in 1 2; do
timeout 600 python bench.py --config sessions --no-cpu-baseline > gpurun_out/r2_bench_sessions_$i.json 2> gpurun_out/r2_bench_sessions_$i.err; echo "sessions rc=$?"
tail -2 gpurun_out/r2_bench_sessions_$i.err | cut -c1-300
python -c "
import json;d=json.load(open('gpurun_out/r2_bench_sessions_$i.json'));print('SESS',d['value'],d['e2e']['p50_ms'],d['rank0_dispatcher'])"
done
