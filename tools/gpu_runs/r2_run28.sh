mkdir -p gpurun_out
nproc
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --no-cpu-baseline > gpurun_out/n4_c512.json 2> gpurun_out/n4_c512.err
tail -2 gpurun_out/n4_c512.err | cut -c1-300
python -c "
import json;d=json.load(open('gpurun_out/n4_c512.json'));print('N4 C512',d['value'],d['e2e']['value'],round(d['e2e']['value']/d['value'],4),d['e2e']['p50_ms'],d['tuning']['table_misses'],d['e2e']['output_matches_golden'],d['roofline']['frac'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 4 --config sessions --no-cpu-baseline > gpurun_out/n4_sessions.json 2> gpurun_out/n4_sessions.err
tail -2 gpurun_out/n4_sessions.err | cut -c1-300
python -c "
import json;d=json.load(open('gpurun_out/n4_sessions.json'));print('N4 SESS',d['value'],d['e2e']['p50_ms'],d['rank0_dispatcher'])"
