mkdir -p gpurun_out
XRT_B200_LIB=variants/checked/libxrt_b200.so timeout 400 python tools/gpu_tile_check.py --hash-only gpurun_out/c18_hash_checked.json > gpurun_out/c18_hash_checked.log 2>&1
timeout 400 python tools/gpu_tile_check.py --hash-only gpurun_out/c18_hash_nr.json > gpurun_out/c18_hash_nr.log 2>&1
timeout 300 python tools/gpu_reflect_time.py --tag final > gpurun_out/c18_time.log 2>&1
timeout 600 python -m pytest tests/test_dcm.py tests/test_materials.py tests/test_multilayer.py tests/test_thin_crystals.py tests/test_laue.py tests/test_fused_chain.py tests/test_planes.py -m gpu -x -q > gpurun_out/c18_pytest.log 2>&1
