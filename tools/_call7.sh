mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:reflect_tile_kernel<.int.1, .int.1, .bool.0' -c 1 -o gpurun_out/prof_tile_nr python tools/gpu_reflect_time.py --nrays 50000000 --reps 1 --no-dcm --tag ncu > gpurun_out/c7_ncu1.log 2>&1
XRTB_REFLECT_TILE=0 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:reflect_main_kernel<.int.0, .int.3, .bool.0' -c 2 -o gpurun_out/prof_crystal_old python tools/gpu_reflect_time.py --nrays 50000000 --reps 1 --tag ncu > gpurun_out/c7_ncu2.log 2>&1
timeout 400 python tools/gpu_tile_check.py > gpurun_out/c7_tile_check.log 2>&1
timeout 300 python tools/gpu_reflect_time.py --tag nr3 > gpurun_out/c7_time.log 2>&1
ls -la gpurun_out/*.ncu-rep
