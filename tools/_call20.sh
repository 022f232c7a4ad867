mkdir -p gpurun_out
XRT_B200_LIB=variants/nogc/libxrt_b200.so timeout 300 python tools/gpu_shine_hash.py gpurun_out/c20_hash_nogc.json > gpurun_out/c20_a.log 2>&1
timeout 300 python tools/gpu_shine_hash.py gpurun_out/c20_hash_gc.json > gpurun_out/c20_b.log 2>&1
B="--steps 3 --warmup 2 --no-kirchhoff --no-dcm --no-config5 --no-cpu-baseline"
XRT_B200_LIB=variants/nogc/libxrt_b200.so timeout 300 python bench.py $B > gpurun_out/c20_bench_nogc.json 2> gpurun_out/c20_bench_nogc.err
timeout 300 python bench.py $B > gpurun_out/c20_bench_gc.json 2> gpurun_out/c20_bench_gc.err
timeout 600 python -m pytest tests/test_undulator.py tests/test_shine.py tests/test_shine_wave.py tests/test_fullsize.py -m gpu -x -q > gpurun_out/c20_pytest.log 2>&1
