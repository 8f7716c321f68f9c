mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3200 --csv --log-file gpurun_out/r02_launches.csv python tools/profile_frame.py --frames 2 --eager > gpurun_out/r02_ncu.log 2>&1
tail -3 gpurun_out/r02_ncu.log
timeout 300 python tools/profile_frame.py --frames 3 --sections 2 > gpurun_out/r02_sections2.txt 2>&1
head -5 gpurun_out/r02_sections2.txt
timeout 300 python tools/profile_frame.py --frames 3 --sections 4 > gpurun_out/r02_sections4.txt 2>&1
