#!/bin/bash
# Run under gpurun: launch list + full ncu capture of the top kernels (B200_PROFILING.md recipe), round 2.
set -x
mkdir -p gpurun_out
NR=${NR:-50000000}
COMMON="--no-cpu-baseline --no-rejection --no-dcm --no-config5"
# (1) every launch of two bench steps with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_rays.csv \
    python bench.py --steps 2 --warmup 1 --nrays $NR --no-kirchhoff $COMMON > gpurun_out/bench_under_ncu_rays.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_kirchhoff.csv \
    python bench.py --steps 1 --warmup 1 --nrays 1000000 --kirchhoff-n 200000 --kirchhoff-m 262144 $COMMON > gpurun_out/bench_under_ncu_kir.log 2>&1
# (2) full-set captures of the top kernels (first launch of each after the warm-up step)
ncu --set full --clock-control none --import-source on -k regex:'shine_uniform_kernel|reflect_main_kernel|reflect_tile_kernel|reflect_prepass|expose_kernel' -s 6 -c 8 \
    -o gpurun_out/prof_rays python bench.py --steps 1 --warmup 1 --nrays $NR --no-kirchhoff $COMMON > gpurun_out/prof_rays.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'kirchhoff_kernel' -c 1 \
    -o gpurun_out/prof_kirchhoff python bench.py --steps 1 --warmup 1 --nrays 1000000 --kirchhoff-n 100000 --kirchhoff-m 262144 $COMMON > gpurun_out/prof_kir.log 2>&1
ls -la gpurun_out
