mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kirchhoff.py tests/test_math_helpers.py tests/test_waves.py tests/test_waves_oe.py tests/test_pipeline.py tests/test_abi.py -m gpu -x -q > gpurun_out/c10_pytest.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 2 --no-dcm --no-config5 --no-rejection --no-cpu-baseline > gpurun_out/c10_bench.json 2> gpurun_out/c10_bench.err
