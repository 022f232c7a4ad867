mkdir -p gpurun_out
XRT_B200_LIB=variants/pref/libxrt_b200.so timeout 300 python tools/gpu_reflect_time.py --tag pref --no-dcm > gpurun_out/c22_time_pref.log 2>&1
timeout 300 python tools/gpu_reflect_time.py --tag base --no-dcm > gpurun_out/c22_time_base.log 2>&1
