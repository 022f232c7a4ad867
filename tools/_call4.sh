mkdir -p gpurun_out
for T in t256 t128; do
  XRT_B200_LIB=variants/$T/libxrt_b200.so timeout 300 python tools/gpu_reflect_time.py --tag $T > gpurun_out/c4_time_$T.log 2>&1
done
timeout 300 python tools/gpu_reflect_time.py --tag t512 > gpurun_out/c4_time_t512.log 2>&1
