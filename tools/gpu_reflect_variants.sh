#!/bin/bash
# GPU-side experiment: reflect_main_kernel variants (occupancy via __launch_bounds__ min blocks, L2 prefetch of the
# field arrays).  VARIANTS: space-separated "MINBLOCKS:extra-nvcc-flag" entries.
mkdir -p gpurun_out
for V in ${VARIANTS:-"4:" "4:-DXRTB_REFLECT_NO_PREFETCH" "5:" "3:"}; do
  MB=${V%%:*}; EX=${V#*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -fmad=false -DXRTB_REFLECT_MINBLOCKS=$MB $EX \
     -Xptxas -v -I include -c -o build/obj/reflect.o xrt_b200/csrc/reflect.cu 2>&1 | grep -A2 "reflect_main_kernelILi1ELi1E" | grep -E "registers|spill" | head -2
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o xrt_b200/libxrt_b200.so build/obj/*.o
  echo "VARIANT MINBLOCKS=$MB $EX"
  python bench.py --steps 3 --warmup 3 --nrays 100000000 --no-kirchhoff --no-cpu-baseline --no-rejection --no-dcm 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value', d['value'], 'ms', d['ms_per_step'], {k: round(v['ms'],2) for k,v in d['stages'].items()})
"
done
