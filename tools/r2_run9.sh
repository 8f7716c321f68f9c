set -x
mkdir -p gpurun_out
for poly in 4 0 3 2; do
  VSD_ATTN_POLY=$poly timeout 300 python tools/gpu_check.py attn_timing 2>&1 | grep "TIME\|EXC" | cut -c1-200
done
VSD_ATTN_V2=0 timeout 300 python tools/gpu_check.py attn_timing 2>&1 | grep "TIME\|EXC" | cut -c1-200
for c in "2 2 48 129 300" "2 2 40 129 300" "2 2 48 256 384" "1 2 48 129 300" "2 8 48 129 300" "2 2 48 129 256" "2 2 48 128 300" "1 3 8 520 520" "4 8 40 2304 2304"; do
  timeout 120 python tools/attn_one.py $c 2>&1 | grep "CASE\|rror" | cut -c1-200
done
VSD_ATTN_V2=0 timeout 300 compute-sanitizer --tool memcheck python tools/attn_one.py 2 2 48 129 300 2>&1 | grep -v "^$" | head -40
export VSD_TUNING_TABLES=0
timeout 600 python tools/gpu_pipeline_check.py 512x512x1 2>&1 | grep "TIMING\|launches\|PSNR\|free-running"
VSD_FF_OUT_FUSE=0 timeout 600 python tools/gpu_pipeline_check.py 512x512x1 2>&1 | grep "TIMING\|launches\|PSNR"
for lanes in 4 8; do
  timeout 900 python bench.py --lanes $lanes --steps 40 --no-cpu-baseline --paced-frames 0 > gpurun_out/r2_bench_l$lanes.json 2> gpurun_out/r2_bench_l$lanes.err
  python -c "
import json;d=json.load(open('gpurun_out/r2_bench_l$lanes.json'));print('LANES',$lanes,d['value'],d['e2e']['value'],d['e2e']['p50_ms'],d['single_lane']['value'],d['tuning']['table_misses'],d['launches_per_frame'])"
done
