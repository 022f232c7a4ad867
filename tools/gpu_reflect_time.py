"""GPU-side: per-kernel CUDA-event times of OE.reflect (config 2u toroid + Au) and DCM.double_reflect (config 3).

    [XRT_B200_LIB=variants/TAG/libxrt_b200.so] python tools/gpu_reflect_time.py [--nrays 100000000] [--tag NAME]

Prints one JSON line; appends it to gpurun_out/reflect_times.jsonl.  Run under gpurun."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from xrt_b200 import _lib  # noqa: E402
from xrt_b200.beamline import BeamLine  # noqa: E402
from xrt_b200.beams import Beam  # noqa: E402
from xrt_b200.materials import Material, CrystalSi  # noqa: E402
from xrt_b200.oes import ToroidMirror, DCM  # noqa: E402
from xrt_b200.planes import Screen  # noqa: E402
from xrt_b200.sources import Undulator, GeometricSource  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--nrays', type=int, default=100000000)
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--tag', default=os.environ.get('XRT_B200_LIB', 'default'))
    ap.add_argument('--no-dcm', action='store_true')
    args = ap.parse_args()
    _lib.init(0)
    torch.cuda.set_device(0)
    n = args.nrays
    bl = BeamLine()
    src = Undulator(bl, 'u', eMin=bench.E1 * 0.995, eMax=bench.E1 * 1.005, seed=1234, **bench.undulator_kwargs(n))
    mirror = ToroidMirror(bl, 'tm', material=Material('Au', rho=19.3), **bench.MIRROR)
    screen = Screen(bl, 'scr', center=(0, bench.SCREEN_Y, 25000. * np.tan(2 * 3e-3)))
    src.reset()
    beam = Beam(n, withAmplitudes=True, uninitialized=True)
    src.shine_uniform_fused(beam, index_offset=0)
    res = {'tag': args.tag, 'nrays': n}
    for needLocal in (True, False):
        mirror.profileEvents = []
        for _ in range(args.reps + 2):
            out = mirror.reflect(beam, needLocal=needLocal, checkTables=False)
            del out
        torch.cuda.synchronize()
        ev = mirror.profileEvents[2:]
        ms = [float(np.median([e[a].elapsed_time(e[b]) for e in ev])) for a, b in ((0, 1), (4, 2), (5, 3))]
        mirror.profileEvents = None
        key = 'toroid_Au_lb' if needLocal else 'toroid_Au_nolb'
        bpr = 404 if needLocal else 264
        res[key] = {'prepassA_ms': ms[0], 'prepassB_ms': ms[1], 'main_ms': ms[2],
                    'main_GBs': bpr * n / ms[2] / 1e6, 'main_frac_hbm': bpr * n / ms[2] / 1e6 / 6545.9}
    gb, lb = mirror.reflect(beam, needLocal=True, checkTables=False)
    res['good_fraction'] = float((lb.state == 1).sum()) / n
    # checksum of the result (variant runs must agree bit for bit)
    res['checksum'] = [float(gb.x.sum()), float(gb.Jss.sum()), float(lb.Es.real.sum()), float(gb.Ep.imag.sum()),
                       int(gb.state.sum())]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(args.reps):
        e0.record()
        s = screen.expose(gb)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
        del s
    res['expose_ms'] = float(np.median(ts))
    # fused reflect + expose (xrt_reflect_expose_f64): main kernel time with and without the local beam
    for needLocal in (True, False):
        mirror.profileEvents = []
        for _ in range(args.reps + 2):
            out = mirror.reflect(beam, needLocal=needLocal, checkTables=False, exposeOn=screen)
            del out
        torch.cuda.synchronize()
        ev = mirror.profileEvents[2:]
        res['fused_expose_' + ('lb' if needLocal else 'nolb')] = {
            'main_ms': float(np.median([e[5].elapsed_time(e[3]) for e in ev]))}
        mirror.profileEvents = None
    del gb, lb, beam
    if not args.no_dcm:
        bl3 = BeamLine()
        gs = GeometricSource(bl3, 'gs', nrays=n, dx=0.2, dz=0.02, dxprime=5e-5, dzprime=2e-5, distE='flat',
                             energies=(8995., 9005.), polarization='h', seed=4242)
        si = CrystalSi(hkl=(1, 1, 1))
        E0 = np.array([9000.])
        theta = float(si.get_Bragg_angle(E0)[0] - si.get_dtheta(E0)[0])
        dcm = DCM(bl3, 'dcm', center=(0, 2e4, 0), bragg=theta, material=si, material2=si, cryst2perpTransl=10.,
                  limPhysX=[-50, 50], limPhysY=[-100, 100], limPhysX2=[-50, 50], limPhysY2=[-100, 200])
        b = gs.shine(withAmplitudes=True)
        ts = []
        for _ in range(args.reps + 1):
            e0.record()
            out = dcm.double_reflect(b, needLocal=True, checkTables=False)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
            chk = [float(out[0].x.sum()), float(out[0].Jss.sum()), int(out[0].state.sum())]
            del out
        res['dcm_double_reflect_ms'] = float(np.median(ts[1:]))
        res['dcm_checksum'] = chk
    line = json.dumps(res)
    print(line)
    os.makedirs('gpurun_out', exist_ok=True)
    with open('gpurun_out/reflect_times.jsonl', 'a') as f:
        f.write(line + '\n')


if __name__ == '__main__':
    main()
