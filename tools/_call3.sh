set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:reflect_tile_kernel -c 3 -o gpurun_out/prof_tile python tools/gpu_reflect_time.py --nrays 50000000 --reps 1 --no-dcm --tag ncu > gpurun_out/c3_ncu.log 2>&1
