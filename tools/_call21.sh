mkdir -p gpurun_out
timeout 300 python tools/gpu_shine_hash.py gpurun_out/c21_hash.json > gpurun_out/c21.log 2>&1
timeout 600 python -m pytest tests/test_undulator.py tests/test_shine.py tests/test_synchrotron.py tests/test_dropin_reference.py -m gpu -x -q > gpurun_out/c21_pytest.log 2>&1
timeout 300 python bench.py --steps 3 --warmup 2 --no-kirchhoff --no-dcm --no-config5 --no-cpu-baseline > gpurun_out/c21_bench.json 2> gpurun_out/c21_bench.err
