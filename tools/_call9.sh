mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/c9_bench.json 2> gpurun_out/c9_bench.err
timeout 300 python tools/gpu_reflect_time.py --tag preB > gpurun_out/c9_time.log 2>&1
