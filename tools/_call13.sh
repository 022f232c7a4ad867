mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_math_helpers.py tests/test_reflect_tile.py -m gpu -x -q > gpurun_out/c13_pytest.log 2>&1
