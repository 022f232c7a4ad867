"""GPU-side exploration: undulator kernel register blocking (run under gpurun)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xrt_b200 import kernels, _lib
from xrt_b200.sources import Undulator
from xrt_b200.beamline import BeamLine
import bench
_lib.init(0)
n = 50000000
for Kx in (0., 0.5):
    kw = bench.undulator_kwargs(n)
    if Kx:
        kw.update(K=0, Kx=Kx, Ky=1.45)
    src = Undulator(BeamLine(), 'u', eMin=bench.E1*0.995, eMax=bench.E1*1.005, **kw)
    src.reset()
    s = src.draw_samples(n)
    for R in (1, 2, 4):
        ts = []
        for it in range(3):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            kernels.undulator_imap(src.Kx, src.Ky, src.L0, src.Np, src.gamma, 1.0, src.dstep, 0, s['E'], s['theta'], s['psi'], src._tables(s['E'].device), R=R)
            e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        t = min(ts)
        print('Kx=%g R=%d  %.2f ms  %.3e sample-node/s  %.2f TFLOP/s(63)' % (Kx, R, t, n*64/t*1e3, 63*n*64/t*1e3/1e12), flush=True)
