mkdir -p gpurun_out
B="--steps 3 --warmup 2 --no-kirchhoff --no-dcm --no-config5 --no-rejection --no-cpu-baseline"
XRT_B200_LIB=variants/un2/libxrt_b200.so timeout 300 python bench.py $B > gpurun_out/c8_bench_un2.json 2> gpurun_out/c8_bench_un2.err
timeout 300 python bench.py $B > gpurun_out/c8_bench_def.json 2> gpurun_out/c8_bench_def.err
timeout 300 python tools/gpu_reflect_time.py --tag nr4 > gpurun_out/c8_time.log 2>&1
