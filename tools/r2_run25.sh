mkdir -p gpurun_out
for L in 3 4; do
timeout 600 python bench.py --config sessions --session-lanes $L --no-cpu-baseline > gpurun_out/s_l$L.json 2> gpurun_out/s_l$L.err; echo "sessions lanes $L rc=$?"
tail -2 gpurun_out/s_l$L.err | cut -c1-300
python -c "
import json;d=json.load(open('gpurun_out/s_l$L.json'));print('SESS',d['value'],d['e2e']['p50_ms'],d['rank0_dispatcher'])"
done
timeout 900 python bench.py --config c768b4 --lanes-c768 2 --no-cpu-baseline > gpurun_out/c768_l2.json 2> gpurun_out/c768_l2.err
tail -2 gpurun_out/c768_l2.err | cut -c1-300
python -c "
import json;d=json.load(open('gpurun_out/c768_l2.json'));print('C768 lanes2',d['value'],d['e2e']['value'],d['e2e']['p50_ms'],d['tuning']['table_misses'],d['roofline']['frac'])"
timeout 900 python bench.py --lanes 8 --paced-frames 0 --no-cpu-baseline > gpurun_out/c512_l8.json 2> gpurun_out/c512_l8.err
python -c "
import json;d=json.load(open('gpurun_out/c512_l8.json'));print('C512 lanes8',d['value'],d['e2e']['value'],d['e2e']['p50_ms'],d['tuning']['table_misses'],d['roofline']['frac'])"
