mkdir -p gpurun_out
XRT_B200_LIB=variants/srcchk/libxrt_b200.so timeout 300 python tools/gpu_shine_hash.py gpurun_out/c17_hash_checked.json > gpurun_out/c17_a.log 2>&1
timeout 300 python tools/gpu_shine_hash.py gpurun_out/c17_hash_nr.json > gpurun_out/c17_b.log 2>&1
B="--steps 3 --warmup 2 --no-kirchhoff --no-dcm --no-config5 --no-cpu-baseline"
XRT_B200_LIB=variants/srcchk/libxrt_b200.so timeout 300 python bench.py $B > gpurun_out/c17_bench_checked.json 2> gpurun_out/c17_bench_checked.err
timeout 300 python bench.py $B > gpurun_out/c17_bench_nr.json 2> gpurun_out/c17_bench_nr.err
timeout 600 python -m pytest tests/test_undulator.py tests/test_shine.py tests/test_shine_wave.py tests/test_synchrotron.py -m gpu -x -q > gpurun_out/c17_pytest.log 2>&1
