mkdir -p gpurun_out
timeout 300 python tools/gpu_reflect_time.py --tag pref2 > gpurun_out/c23_time.log 2>&1
timeout 400 python tools/gpu_tile_check.py --hash-only gpurun_out/c23_hash.json > gpurun_out/c23_hash.log 2>&1
timeout 600 python -m pytest tests/test_reflect.py tests/test_dcm.py tests/test_multiple_reflect.py tests/test_gratings.py tests/test_refractive.py tests/test_fullsize.py -m gpu -x -q > gpurun_out/c23_pytest.log 2>&1
