set -x
mkdir -p gpurun_out
for poly in 4 0 3 2; do
  VSD_ATTN_POLY=$poly timeout 300 python tools/gpu_check.py attn > gpurun_out/r2_attn_poly$poly.txt 2>&1
  grep -c PASS gpurun_out/r2_attn_poly$poly.txt; grep "FAIL\|EXC\|DONE\|TIME" gpurun_out/r2_attn_poly$poly.txt | cut -c1-300 | head -30
done
VSD_ATTN_V2=0 timeout 300 python tools/gpu_check.py attn > gpurun_out/r2_attn_v1.txt 2>&1
grep "FAIL\|EXC\|DONE\|TIME" gpurun_out/r2_attn_v1.txt | cut -c1-300 | head
for lanes in 4 6; do
  timeout 900 python bench.py --lanes $lanes --steps 40 --no-cpu-baseline --paced-frames 0 > gpurun_out/r2_bench_l$lanes.json 2> gpurun_out/r2_bench_l$lanes.err
  python -c "
import json;d=json.load(open('gpurun_out/r2_bench_l$lanes.json'));print('LANES',$lanes,d['value'],d['e2e']['value'],d['single_lane']['value'],d['tuning'],d['launches_per_frame'])"
done
