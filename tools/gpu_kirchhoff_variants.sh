#!/bin/bash
# GPU-side experiment: kirchhoff_kernel occupancy variants (__launch_bounds__ min blocks / CTA size).
# VARIANTS: space-separated "THREADS:MINBLOCKS" entries.  The library is rebuilt on the box for each.
mkdir -p gpurun_out build/obj
for f in capi rays reflect source undulator undulator_ext; do
  [ -f build/obj/$f.o ] || { echo "build/obj/$f.o missing: ship build/ (not in .gpurunignore) for this experiment"; exit 1; }
done
for V in ${VARIANTS:-"128:2" "128:3" "256:1" "64:6"}; do
  TH=${V%%:*}; MB=${V#*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DXRTB_KIRCHHOFF_MINBLOCKS=$MB -DXRTB_KIRCHHOFF_THREADS=$TH \
     -Xptxas -v -I include -c -o build/obj/kirchhoff.o xrt_b200/csrc/kirchhoff.cu 2>&1 | grep -A2 "kirchhoff_kernelILi4ELi${TH}ELi128ELb1ELb1" | grep -E "registers|spill" | head -2
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o xrt_b200/libxrt_b200.so build/obj/*.o
  echo "VARIANT THREADS=$TH MINBLOCKS=$MB"
  python bench.py --steps 1 --warmup 1 --nrays 1000000 --kirchhoff-n ${KN:-400000} --kirchhoff-m ${KM:-1048576} --no-cpu-baseline --no-rejection --no-dcm 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)['kirchhoff']; print('ms', d['ms'], 'pairs/s', d['pairs_per_s'], 'frac', d['roofline']['frac'], 'issue', d['roofline']['issue_slot_frac'])
"
done
