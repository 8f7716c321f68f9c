mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/n2_c512.json 2> gpurun_out/n2_c512.err
tail -2 gpurun_out/n2_c512.err | cut -c1-300
python -c "
import json;d=json.load(open('gpurun_out/n2_c512.json'));print('N2 C512',d['value'],d['e2e']['value'],d['e2e']['p50_ms'],d['tuning']['table_misses'],d['e2e']['output_matches_golden'],d['roofline']['frac'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --config sessions --no-cpu-baseline > gpurun_out/n2_sessions.json 2> gpurun_out/n2_sessions.err
tail -2 gpurun_out/n2_sessions.err | cut -c1-300
python -c "
import json;d=json.load(open('gpurun_out/n2_sessions.json'));print('N2 SESS',d['value'],d['e2e']['p50_ms'],d['rank0_dispatcher'])"
