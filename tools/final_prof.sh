set -x
rm -f gpurun_out/tune_cache_512x512x1.txt
python tools/profile_frame.py --frames 3 --save-tuning > gpurun_out/m_tune.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4200 --csv --log-file gpurun_out/m_launches.csv python tools/profile_frame.py --frames 2 --eager > gpurun_out/m_ncu.log 2>&1
python tools/profile_frame.py --frames 2 --sections 2 > gpurun_out/m_sections2.txt 2>&1
tail -3 gpurun_out/m_ncu.log
