mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanity_smoke.log 2>&1
timeout 300 python bench.py --steps 3 --warmup 3 --no-kirchhoff --no-dcm --no-config5 --no-rejection --no-cpu-baseline > gpurun_out/sanity_bench.json 2> gpurun_out/sanity_bench.err
timeout 600 python -m pytest tests/test_fused_chain.py tests/test_reflect_tile.py tests/test_shine.py tests/test_abi.py -m gpu -x -q > gpurun_out/sanity_pytest.log 2>&1
