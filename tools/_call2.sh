set -x
mkdir -p gpurun_out
timeout 600 python tools/gpu_tile_check.py > gpurun_out/c2_tile_check.log 2>&1
echo "rc=$?" >> gpurun_out/c2_tile_check.log
XRTB_REFLECT_TILE=0 timeout 300 python tools/gpu_reflect_time.py --tag old_kernel > gpurun_out/c2_time_old.log 2>&1
XRTB_REFLECT_TILE=1 timeout 300 python tools/gpu_reflect_time.py --tag tile_kernel > gpurun_out/c2_time_tile.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 1 --no-kirchhoff --no-rejection --no-cpu-baseline --config5-wave-n 20000 --config5-wave-m 65536 > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err
