mkdir -p gpurun_out
for m in auto 0 auto; do
if [ $m = auto ]; then unset VSD_BLOCKING_SYNC; else export VSD_BLOCKING_SYNC=$m; fi
timeout 600 python bench.py --no-cpu-baseline --steps 240 > gpurun_out/ab2_sync$m.json 2> gpurun_out/ab2_sync$m.err
tail -1 gpurun_out/ab2_sync$m.err | cut -c1-200
python -c "
import json;d=json.load(open('gpurun_out/ab2_sync$m.json'));print('SYNC=$m value',round(d['value'],2),'e2e',round(d['e2e']['value'],2),'ratio',round(d['e2e']['value']/d['value'],4),'p50',round(d['e2e']['p50_ms'],2),'single e2e/value',round(d['single_lane']['e2e'],2),round(d['single_lane']['value'],2),'paced',d['paced_30fps']['p50_ms'],d['paced_30fps']['p95_ms'], d['e2e']['output_matches_golden'])"
done
unset VSD_BLOCKING_SYNC
timeout 600 python bench.py --config sessions --no-cpu-baseline > gpurun_out/ab2_sessions.json 2> gpurun_out/ab2_sessions.err
python -c "
import json;d=json.load(open('gpurun_out/ab2_sessions.json'));print('SESS',d['value'],d['e2e']['p50_ms'])"
timeout 900 python -m pytest tests/test_gpu_lanes.py -q -m gpu -x 2>&1 | tail -3
