mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kirchhoff.py tests/test_waves.py tests/test_waves_oe.py tests/test_pipeline.py tests/test_dropin_reference.py -m gpu -x -q > gpurun_out/c19_pytest.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 2 --no-dcm --no-config5 --no-rejection --no-cpu-baseline > gpurun_out/c19_bench.json 2> gpurun_out/c19_bench.err
