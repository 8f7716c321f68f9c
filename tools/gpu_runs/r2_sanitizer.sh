# compute-sanitizer over one eager frame at the odd 360x640 geometry with the strided-TMA downsampling forced on (the path
# whose round-1 NaN was never named), then initcheck (uninitialised global reads) over the same frame.
mkdir -p gpurun_out
export VSD_TMA_S2=1
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/profile_frame.py --frames 1 --eager --size 360x640x1 > gpurun_out/r02_sanitizer_memcheck.txt 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|launches/frame|Invalid|SETUP_DONE" gpurun_out/r02_sanitizer_memcheck.txt | head
timeout 900 compute-sanitizer --tool initcheck --print-limit 20 python tools/profile_frame.py --frames 1 --eager --size 360x640x1 > gpurun_out/r02_sanitizer_initcheck.txt 2>&1
echo "initcheck rc=$?"; grep -E "ERROR SUMMARY|launches/frame|Uninitialized|SETUP_DONE" gpurun_out/r02_sanitizer_initcheck.txt | head
