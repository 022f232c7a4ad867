"""Time the one-pass plot histogram kernel on a synthetic 1e8-ray screen beam (run under gpurun)."""
import sys
import torch
sys.path.insert(0, '.')
from xrt_b200 import histogram, _lib  # noqa: E402
from xrt_b200.beams import Beam  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100000000
_lib.init()
torch.manual_seed(3)
b = Beam(n, device='cuda', uninitialized=True)
b.x = torch.randn(n, dtype=torch.float64, device='cuda') * 2.0
b.z = torch.randn(n, dtype=torch.float64, device='cuda') * 0.5
b.E = torch.rand(n, dtype=torch.float64, device='cuda') * 100 + 8950
b.Jss = torch.rand(n, dtype=torch.float64, device='cuda')
b.Jpp = torch.rand(n, dtype=torch.float64, device='cuda') * 0.1
b.Jsp = torch.zeros(n, dtype=torch.complex128, device='cuda')
b.state = torch.ones(n, dtype=torch.int32, device='cuda')
for bins in (128, 256, 1024):
    ph = histogram.PlotHistograms(histogram.Axis('x', [-8, 8], bins), histogram.Axis('z', [-2, 2], bins),
                                  histogram.Axis('E', [8950, 9050], bins))
    ph.accumulate(b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ph.accumulate(b)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print('bins', bins, 'plot_hist %.2f ms' % ms, '%.0f Mrays/s' % (n / ms / 1e3), 'read GB/s %.0f' % (n * 44 / ms / 1e6))
    img = torch.zeros((bins, bins), dtype=torch.float64, device='cuda')
    histogram.hist2d(b, [-8, 8], [-2, 2], bins=(bins, bins), out=img)
    e0.record()
    for _ in range(3):
        histogram.hist2d(b, [-8, 8], [-2, 2], bins=(bins, bins), out=img)
    e1.record()
    torch.cuda.synchronize()
    print('bins', bins, 'hist2d only %.2f ms' % (e0.elapsed_time(e1) / 3))
