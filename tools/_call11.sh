mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_undulator.py tests/test_shine.py tests/test_math_helpers.py tests/test_fullsize.py tests/test_pipeline.py -m gpu -x -q > gpurun_out/c11_pytest.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 2 --no-kirchhoff --no-dcm --no-config5 --no-cpu-baseline > gpurun_out/c11_bench.json 2> gpurun_out/c11_bench.err
