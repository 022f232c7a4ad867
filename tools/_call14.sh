mkdir -p gpurun_out
XRT_B200_LIB=variants/checked/libxrt_b200.so timeout 400 python tools/gpu_tile_check.py --hash-only gpurun_out/c14_hash_checked.json > gpurun_out/c14_hash_checked.log 2>&1
timeout 400 python tools/gpu_tile_check.py --hash-only gpurun_out/c14_hash_nr.json > gpurun_out/c14_hash_nr.log 2>&1
timeout 600 python -m pytest tests/test_reflect.py tests/test_pipeline.py tests/test_waves_oe.py tests/test_oes_more.py -m gpu -x -q > gpurun_out/c14_pytest.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 2 --no-kirchhoff --no-rejection --no-cpu-baseline --config5-wave-n 20000 --config5-wave-m 65536 > gpurun_out/c14_bench.json 2> gpurun_out/c14_bench.err
