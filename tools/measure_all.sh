# One measurement pass on a B200 box (run through gpurun); everything lands in gpurun_out/m_*.
set -x
mkdir -p gpurun_out
rm -f gpurun_out/tune_cache_512x512x1.txt
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/m_smoke.log 2>&1
python tools/profile_frame.py --frames 3 --save-tuning > gpurun_out/m_tune.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4200 --csv --log-file gpurun_out/m_launches.csv python tools/profile_frame.py --frames 2 --eager > gpurun_out/m_ncu.log 2>&1
python bench.py --steps 30 --warmup 5 > gpurun_out/m_bench.json 2> gpurun_out/m_bench.err
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/m_bench_ref.json 2> gpurun_out/m_bench_ref.err
for n in 1 2 3 4; do python tools/concurrent_lanes.py $n 2>&1 | tail -1; done > gpurun_out/m_lanes.txt
python tools/gpu_pipeline_check.py 768x768x4 2>&1 | tail -4 > gpurun_out/m_768.txt
python tools/gpu_pipeline_check.py 512x512x1 --controlnet 2>&1 | tail -4 > gpurun_out/m_cn.txt
python tools/gpu_pipeline_check.py 512x512x1 --kl 2>&1 | tail -4 > gpurun_out/m_kl.txt
python tools/multi_session_sim.py --streams 8 --batch 4 --lanes 2 --frames 60 --switch-every 20 2>&1 | tail -1 > gpurun_out/m_sessions.txt
python tools/profile_frame.py --frames 2 --sections 2 > gpurun_out/m_sections2.txt 2>&1
python tools/gpu_check.py clip resize 2>&1 | tail -6 > gpurun_out/m_clip_resize.txt
cat gpurun_out/m_lanes.txt gpurun_out/m_768.txt gpurun_out/m_cn.txt gpurun_out/m_kl.txt gpurun_out/m_sessions.txt | cut -c1-300
tail -2 gpurun_out/m_smoke.log
cut -c1-300 gpurun_out/m_bench.json
