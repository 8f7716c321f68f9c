# Round-2 final measurements: smoke, GPU tests, the bench configurations, ncu launch list / section timings / one --set full capture.
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
timeout 1800 python -m pytest tests -q -m gpu -x -s > gpurun_out/r02_pytest_gpu.txt 2>&1
grep "AutoencoderKL per-step" gpurun_out/r02_pytest_gpu.txt; tail -3 gpurun_out/r02_pytest_gpu.txt
timeout 900 python bench.py > gpurun_out/r02_bench_c512.json 2> gpurun_out/r02_bench_c512.err
tail -2 gpurun_out/r02_bench_c512.err | cut -c1-300
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_c512.json'));print('C512',d['value'],d['e2e']['value'],d['e2e']['p50_ms'],d['single_lane']['value'],d['tuning']['table_misses'],d['launches_per_frame'],d['e2e']['output_matches_golden'],d['roofline'],d.get('paced_30fps'),d.get('cpu_baseline'),d['clocks'])"
timeout 900 python bench.py --config c768b4 --no-cpu-baseline > gpurun_out/r02_bench_c768b4.json 2> gpurun_out/r02_bench_c768b4.err
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_c768b4.json'));print('C768',d['value'],d['e2e']['value'],d['e2e']['p50_ms'],d['tuning']['table_misses'],d['e2e'].get('output_matches_golden'),d['roofline']['frac'])"
timeout 600 python bench.py --config sessions --no-cpu-baseline > gpurun_out/r02_bench_sessions.json 2> gpurun_out/r02_bench_sessions.err
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_sessions.json'));print('SESS',d['value'],d['e2e']['p50_ms'],d['rank0_dispatcher'],d['roofline']['frac'])"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3200 --csv --log-file gpurun_out/r02_launches_final.csv python tools/profile_frame.py --frames 2 --eager > gpurun_out/r02_ncu.log 2>&1
tail -2 gpurun_out/r02_ncu.log
timeout 300 python tools/profile_frame.py --frames 3 --sections 2 > gpurun_out/r02_sections2_final.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 2 -c 1 -f -o gpurun_out/r02_prof_gemm_4096x320x2880_final python tools/ncu_one_gemm.py 4096 320 320 9 96 1 769 1 > gpurun_out/ncu_g1.log 2>&1
tail -2 gpurun_out/ncu_g1.log
