mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_figure_error.py tests/test_spliced.py tests/test_reflect.py tests/test_multiple_reflect.py -m gpu -x -q > gpurun_out/c16_pytest.log 2>&1
