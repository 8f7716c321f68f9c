mkdir -p gpurun_out
nproc
for m in 0 1 0 1; do
VSD_BLOCKING_SYNC=$m timeout 600 python bench.py --no-cpu-baseline --paced-frames 0 --steps 240 > gpurun_out/ab_sync$m.json 2> gpurun_out/ab_sync$m.err
python -c "
import json;d=json.load(open('gpurun_out/ab_sync$m.json'));print('BLOCKING=$m value',round(d['value'],2),'e2e',round(d['e2e']['value'],2),'ratio',round(d['e2e']['value']/d['value'],4),'p50',round(d['e2e']['p50_ms'],2),'single e2e/value',round(d['single_lane']['e2e'],2),round(d['single_lane']['value'],2))"
done
