mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multiple_reflect.py tests/test_spliced.py tests/test_abi.py tests/test_figure_error.py -m gpu -x -q > gpurun_out/c15_pytest.log 2>&1
