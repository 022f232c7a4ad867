#!/bin/bash
# GPU-side experiment: reflect.cu built with different -D switches (one per variant), timed with bench.py.
# DEFS: ';'-separated variants, each a (possibly empty) list of nvcc -D flags.  build/obj must ship (empty
# .gpurunignore for this call): only reflect.o is recompiled on the box.
mkdir -p gpurun_out
IFS=';' read -ra VARS <<< "${DEFS:-;-DXRTB_HOIST_TABLE_SEARCH}"
for V in "${VARS[@]}"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -fmad=false $V \
     -Xptxas -v -I include -c -o build/obj/reflect.o xrt_b200/csrc/reflect.cu 2>&1 | grep -A2 "reflect_main_kernelILi1ELi1E" | grep -E "registers|spill" | head -2
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o xrt_b200/libxrt_b200.so build/obj/*.o
  echo "VARIANT [$V]"
  python bench.py --steps 5 --warmup 3 --nrays 100000000 --no-kirchhoff --no-cpu-baseline --no-rejection ${BENCH_EXTRA:---no-dcm} 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value', d['value'], 'ms', d['ms_per_step'], {k: round(v['ms'],2) for k,v in d['stages'].items()}, d.get('dcm', {}).get('ms_per_step'))
"
done
