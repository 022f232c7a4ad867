"""GPU-side: fp64 ceilings (DFMA / DMMA / interleaved) + Kirchhoff kernel variants + shard-size scaling (run under gpurun).

Writes gpurun_out/kirchhoff_sweep.json (copied to profiles/r02_kirchhoff_variants.json when judged)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from xrt_b200 import kernels, synthetic, _lib  # noqa: E402

_lib.init(0)
res = {'what': 'B200, fp64; pairs/s and TFLOP/s on the 57-flop-per-pair basis of SURVEY 8d'}
for mode, name in ((0, 'dfma'), (1, 'dmma_m8n8k4'), (2, 'dfma+dmma_interleaved')):
    tf, ms = kernels.fp64_probe(mode, 8192)
    res['probe_' + name] = dict(tflops=tf, ms=ms)
    print('probe', name, tf, 'TFLOP/s', ms, 'ms', flush=True)


def dev(case, sl=None):
    d = {}
    for key in ('xM', 'yM', 'zM'):
        v = case[key] if sl is None else case[key][sl]
        d[key] = torch.from_numpy(np.ascontiguousarray(v)).cuda()
    for key in ('nl', 'k', 'Es', 'Ep'):
        d[key] = torch.from_numpy(np.ascontiguousarray(case[key])).cuda()
    d['pos'] = torch.from_numpy(np.ascontiguousarray(case['pos'].T)).cuda()
    d['nrm'] = torch.from_numpy(np.ascontiguousarray(case['nrm'].T)).cuda()
    return d


def run(d, reps=3, **kw):
    M = d['xM'].numel()
    out = torch.empty((5, M), dtype=torch.complex128, device='cuda')
    times = []
    for it in range(reps + 1):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        kernels.kirchhoff(d['xM'], d['yM'], d['zM'], d['nl'], d['Es'], d['Ep'], d['k'], d['pos'], d['nrm'], out=out,
                          **kw)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    return min(times[1:]), out


# ---- variants at 2e5 x 262144 (the slit of config 4: one common normal)
N, M = 200000, 262144
case = synthetic.kirchhoff_case(7, N, M, mesh=True)
d = dev(case)
ref = None
for name, kw in (('r01_kernel: R4 strict uniform-normal sincos_big', dict(R=4, uniform_normal=True, table_sincos=False)),
                 ('R4 strict uniform-normal sincos_tab (64-entry table)', dict(R=4, uniform_normal=True, table_sincos=True)),
                 ('R4 strict general-normal sincos_tab', dict(R=4, uniform_normal=False, table_sincos=True)),
                 ('R2 strict uniform-normal sincos_tab', dict(R=2, uniform_normal=True, table_sincos=True)),
                 ('R4 RELAXED (FMA-contracted r^2 and phase: not numpy-exact) sincos_tab',
                  dict(R=4, uniform_normal=True, table_sincos=True, relaxed=True))):
    ms, out = run(d, **kw)
    pairs = N * M / (ms * 1e-3)
    entry = dict(ms=ms, pairs_per_s=pairs, tflops57=57 * pairs * 1e-12)
    if ref is None:
        ref = out.clone()
    else:
        entry['max_abs_diff_vs_r01_over_max'] = float((out - ref).abs().max() / ref.abs().max())
    res[name] = entry
    print(name, '%.2f ms  %.3e pairs/s  %.2f TFLOP/s(57)' % (ms, pairs, 57 * pairs * 1e-12), entry.get(
        'max_abs_diff_vs_r01_over_max'), flush=True)

# ---- strong scaling of config 4 on one GPU: the shard sizes ranks get at 1 / 2 / 4 / 8 GPUs (N = 1e6 samples)
if '--no-scaling' not in sys.argv:
    N, Mfull = 1000000, 1024 * 1024
    case = synthetic.kirchhoff_case(7, N, Mfull, mesh=True)
    lib = _lib.load()
    t1 = None
    for world in (8, 4, 2, 1):
        Mloc = Mfull // world
        d = dev(case, slice(0, Mloc))
        ms, _ = run(d, reps=1, R=4, uniform_normal=True)
        ns = None
        if world == 1:
            t1 = ms
        res['config4_shard_of_%d' % world] = dict(M=Mloc, ms=ms, pairs_per_s=N * Mloc / (ms * 1e-3))
        print('shard of', world, Mloc, '%.1f ms' % ms, flush=True)
    for world in (2, 4, 8):
        res['config4_shard_of_%d' % world]['strong_scaling_efficiency_kernel_only'] = \
            t1 / world / res['config4_shard_of_%d' % world]['ms']
os.makedirs('gpurun_out', exist_ok=True)
json.dump(res, open('gpurun_out/kirchhoff_sweep.json', 'w'), indent=1)
print(json.dumps({k: v for k, v in res.items() if k.startswith('config4')}))
