mkdir -p gpurun_out
XRT_B200_LIB=variants/checked/libxrt_b200.so timeout 400 python tools/gpu_tile_check.py --hash-only gpurun_out/c6_hash_checked.json > gpurun_out/c6_hash_checked.log 2>&1
timeout 400 python tools/gpu_tile_check.py --hash-only gpurun_out/c6_hash_nr.json > gpurun_out/c6_hash_nr.log 2>&1
timeout 400 python tools/gpu_tile_check.py > gpurun_out/c6_tile_check.log 2>&1
timeout 300 python tools/gpu_reflect_time.py --tag nr2 > gpurun_out/c6_time_nr.log 2>&1
XRTB_REFLECT_TILE=0 timeout 300 python tools/gpu_reflect_time.py --tag nr2_old > gpurun_out/c6_time_nr_old.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c6_pytest.log 2>&1
