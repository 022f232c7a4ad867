mkdir -p gpurun_out
B="--steps 3 --warmup 2 --no-kirchhoff --no-dcm --no-config5 --no-cpu-baseline"
timeout 600 python -m pytest tests/test_undulator.py tests/test_shine.py -m gpu -x -q > gpurun_out/c12_pytest.log 2>&1
timeout 300 python bench.py $B > gpurun_out/c12_bench_b4.json 2> gpurun_out/c12_bench_b4.err
XRT_B200_LIB=variants/sb1/libxrt_b200.so timeout 300 python bench.py $B > gpurun_out/c12_bench_b1.json 2> gpurun_out/c12_bench_b1.err
XRT_B200_LIB=variants/sb8/libxrt_b200.so timeout 300 python bench.py $B > gpurun_out/c12_bench_b8.json 2> gpurun_out/c12_bench_b8.err
