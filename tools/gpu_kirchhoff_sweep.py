"""GPU-side exploration: fp64 ceilings + Kirchhoff kernel variants (run under gpurun)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from xrt_b200 import kernels, synthetic, _lib

_lib.init(0)
res = {}
for mode, name in ((0, 'dfma'), (1, 'dmma'), (2, 'dfma+dmma')):
    tf, ms = kernels.fp64_probe(mode, 8192)
    res['probe_' + name] = dict(tflops=tf, ms=ms)
    print('probe', name, tf, 'TFLOP/s', ms, 'ms', flush=True)

def dev(case):
    d = {}
    for key in ('xM', 'yM', 'zM', 'nl', 'k'):
        d[key] = torch.from_numpy(np.ascontiguousarray(case[key])).cuda()
    for key in ('Es', 'Ep'):
        d[key] = torch.from_numpy(np.ascontiguousarray(case[key])).cuda()
    d['pos'] = torch.from_numpy(np.ascontiguousarray(case['pos'].T)).cuda()
    d['nrm'] = torch.from_numpy(np.ascontiguousarray(case['nrm'].T)).cuda()
    return d

for (N, M) in ((200000, 262144), (200000, 4096)):
    case = synthetic.kirchhoff_case(1, N, M)
    d = dev(case)
    for R, unroll2 in ((1, False), (1, True), (2, False), (2, True), (4, False)):
        for relaxed in (False,):
            times = []
            for it in range(3):
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                out = kernels.kirchhoff(d['xM'], d['yM'], d['zM'], d['nl'], d['Es'], d['Ep'], d['k'], d['pos'], d['nrm'], R=R, relaxed=relaxed, unroll2=unroll2)
                e1.record(); torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1))
            t = min(times) * 1e-3
            pairs = N * M / t
            key = 'N%d_M%d_R%d_u%d_%s' % (N, M, R, 2 if unroll2 else 1, 'relaxed' if relaxed else 'strict')
            res[key] = dict(ms=t * 1e3, pairs_per_s=pairs, tflops57=57 * pairs * 1e-12)
            print(key, '%.2f ms  %.3e pairs/s  %.2f TFLOP/s(57)' % (t * 1e3, pairs, 57 * pairs * 1e-12), flush=True)
os.makedirs('gpurun_out', exist_ok=True)
json.dump(res, open('gpurun_out/kirchhoff_sweep.json', 'w'), indent=1)
