#!/bin/bash
# GPU-side experiment: reflect_main_kernel variants (block size / occupancy via __launch_bounds__).
# VARIANTS: space-separated "THREADS:MINBLOCKS" entries.
mkdir -p gpurun_out
for V in ${VARIANTS:-"256:4" "128:7" "192:5" "128:6"}; do
  TH=${V%%:*}; MB=${V#*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -fmad=false -DXRTB_REFLECT_MINBLOCKS=$MB -DXRTB_REFLECT_THREADS=$TH \
     -Xptxas -v -I include -c -o build/obj/reflect.o xrt_b200/csrc/reflect.cu 2>&1 | grep -A2 "reflect_main_kernelILi1ELi1E" | grep -E "registers|spill" | head -2
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o xrt_b200/libxrt_b200.so build/obj/*.o
  echo "VARIANT THREADS=$TH MINBLOCKS=$MB"
  python bench.py --steps 3 --warmup 3 --nrays 100000000 --no-kirchhoff --no-cpu-baseline --no-rejection --no-dcm 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value', d['value'], 'ms', d['ms_per_step'], {k: round(v['ms'],2) for k,v in d['stages'].items()})
"
done
