"""GPU-side: reflect_tile_kernel (r02 tile structure of the main pass) against reflect_main_kernel, bit for bit.

    python tools/gpu_tile_check.py [--nrays 1000077]

Every case runs twice in one process -- XRTB_REFLECT_TILE=0 (the old kernel) and =1 -- and every array of every returned
beam plus the reduction slots of the element are compared as raw bits.  Prints one JSON line (appended to
gpurun_out/tile_check.jsonl); exit code 1 on any difference.  Run under gpurun."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from xrt_b200 import _lib, oes  # noqa: E402
from xrt_b200.beamline import BeamLine  # noqa: E402
from xrt_b200.beams import Beam  # noqa: E402
from xrt_b200.materials import Material, CrystalSi  # noqa: E402
from xrt_b200.oes import ToroidMirror, DCM, OE  # noqa: E402
from xrt_b200.planes import Screen  # noqa: E402
from xrt_b200.sources import Undulator, GeometricSource  # noqa: E402


def bits(t):
    if t.dtype == torch.complex128:
        t = torch.view_as_real(t)
    if t.dtype == torch.float64:
        return t.contiguous().view(torch.int64)
    return t


def snapshot(beams, el):
    out = {}
    for bi, b in enumerate(beams):
        if b is None:
            continue
        for k, v in b.tensors().items():
            out['%d.%s' % (bi, k)] = v.clone()
    for key, red in el._red.items():
        out['red'] = red[28:39].clone()  # footprint, counters, sides (not the hint slots: order dependent)
    return out


def compare(a, b):
    bad = []
    for k in a:
        if k not in b:
            bad.append(k + ' missing')
            continue
        x, y = bits(a[k]), bits(b[k])
        if x.shape != y.shape or not torch.equal(x, y):
            nd = int((x != y).sum()) if x.shape == y.shape else -1
            bad.append('%s (%d differ)' % (k, nd))
    return bad


HASH_ONLY = False


def run_both(name, fn, el, results):
    if HASH_ONLY:  # one run with the library as it is: a wrap-around sum of the bit patterns of every array
        beams = fn()
        torch.cuda.synchronize()
        snap = snapshot(beams, el)
        results[name] = {k: int(bits(v).to(torch.int64).sum().item()) for k, v in snap.items()}
        return
    snaps = []
    for flag in ('0', '1'):
        os.environ['XRTB_REFLECT_TILE'] = flag
        beams = fn()
        torch.cuda.synchronize()
        snaps.append(snapshot(beams, el))
        del beams
    bad = compare(snaps[0], snaps[1])
    results[name] = 'ok' if not bad else bad
    print(name, results[name], flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--nrays', type=int, default=1000077)
    ap.add_argument('--hash-only', default='', help='write {case: {array: hash}} of ONE run to this file (compare builds)')
    args = ap.parse_args()
    global HASH_ONLY
    HASH_ONLY = bool(args.hash_only)
    _lib.init(0)
    torch.cuda.set_device(0)
    n = args.nrays
    results = {}
    bl = BeamLine()
    src = Undulator(bl, 'u', eMin=bench.E1 * 0.995, eMax=bench.E1 * 1.005, seed=1234, **bench.undulator_kwargs(n))
    au = Material('Au', rho=19.3)
    mirror = ToroidMirror(bl, 'tm', material=au, **bench.MIRROR)
    screen = Screen(bl, 'scr', center=(0, bench.SCREEN_Y, 25000. * np.tan(2 * 3e-3)))
    src.reset()
    beam = Beam(n, withAmplitudes=True, uninitialized=True)
    src.shine_uniform_fused(beam, index_offset=0)
    # some rays dead on entry, some 'over'
    g = torch.Generator(device='cuda').manual_seed(5)
    rnd = torch.rand(n, device='cuda', generator=g)
    beamD = Beam(copyFrom=beam)
    beamD.state = torch.where(rnd < 0.1, torch.full_like(beam.state, -1), beam.state)
    beamD.state = torch.where((rnd >= 0.1) & (rnd < 0.15), torch.full_like(beam.state, 3), beamD.state)
    beamD.state = torch.where((rnd >= 0.15) & (rnd < 0.2), torch.full_like(beam.state, 2), beamD.state)

    for nm, b in (('toroid', beam), ('toroid_dead', beamD)):
        run_both(nm + '_lb', lambda b=b: mirror.reflect(b, needLocal=True, checkTables=False), mirror, results)
        run_both(nm + '_nolb', lambda b=b: mirror.reflect(b, needLocal=False, checkTables=False)[:1], mirror, results)
        run_both(nm + '_fused_lb', lambda b=b: mirror.reflect(b, needLocal=True, checkTables=False, exposeOn=screen),
                 mirror, results)
        run_both(nm + '_fused_nolb',
                 lambda b=b: mirror.reflect(b, needLocal=False, checkTables=False, exposeOn=screen)[:1], mirror, results)
    # beam without amplitudes, small and ragged sizes, a view at an odd offset (plain-copy path)
    for m in (1, 100, 512, 513, 1023, 4097):
        sub = Beam(copyFrom=beam)
        sub.filter_by_index(torch.arange(m, device='cuda'))
        run_both('toroid_n%d' % m, lambda sub=sub: mirror.reflect(sub, needLocal=True, checkTables=False), mirror,
                 results)
    odd = Beam(n - 1, withAmplitudes=True, uninitialized=True)
    for k, v in beam.tensors().items():
        setattr(odd, k, v[1:])
    run_both('toroid_odd_view', lambda: mirror.reflect(odd, needLocal=True, checkTables=False), mirror, results)
    noamp = Beam(n, withAmplitudes=False, uninitialized=True)
    for k, v in beam.tensors().items():
        if k not in ('Es', 'Ep'):
            setattr(noamp, k, v.clone())
    run_both('toroid_noamp', lambda: mirror.reflect(noamp, needLocal=True, checkTables=False), mirror, results)
    # bare mirror (material kind 0), flat, rolled / yawed, with optical limits
    flat = OE(bl, 'flat', center=(0, 25000, 0), pitch=3e-3, roll=0.2, yaw=1e-3, limPhysX=[-8, 8], limPhysY=[-400, 400],
              limOptX=[-5, 5], limOptY=[-300, 300])
    run_both('flat_bare_lb', lambda: flat.reflect(beamD, needLocal=True, checkTables=False), flat, results)
    flatAu = OE(bl, 'flatAu', center=(0, 25000, 0), pitch=3e-3, positionRoll=np.pi / 2, material=au,
                limPhysX=[-8, 8], limPhysY=[-400, 400])
    run_both('flat_Au_lb', lambda: flatAu.reflect(beamD, needLocal=True, checkTables=False), flatAu, results)
    run_both('flat_Au_fused', lambda: flatAu.reflect(beamD, needLocal=True, checkTables=False, exposeOn=screen), flatAu,
             results)
    # elliptic (parametric) mirror: s, phi, r of the local beam
    g5 = bench.config5_geometry()
    vfm = oes.EllipticalMirrorParam(bl, 'vfm', center=(0, 25000., 0), pitch=3.5e-3, p=25000., q=1800.,
                                    isCylindrical=True, material=au, limPhysX=[-5, 5], limPhysY=[-150, 150])
    run_both('ellipse_lb', lambda: vfm.reflect(beamD, needLocal=True, checkTables=False), vfm, results)
    run_both('ellipse_nolb', lambda: vfm.reflect(beam, needLocal=False, checkTables=False)[:1], vfm, results)
    run_both('ellipse_fused', lambda: vfm.reflect(beam, needLocal=True, checkTables=False, exposeOn=screen), vfm,
             results)
    del g5
    # DCM: two crystal passes, the second with fallback = the incoming beam; small second crystal -> rays lost there
    bl3 = BeamLine()
    gs = GeometricSource(bl3, 'gs', nrays=n, dx=0.2, dz=0.02, dxprime=5e-5, dzprime=2e-5, distE='flat',
                         energies=(8995., 9005.), polarization='h', seed=4242)
    si = CrystalSi(hkl=(1, 1, 1))
    E0 = np.array([9000.])
    theta = float(si.get_Bragg_angle(E0)[0] - si.get_dtheta(E0)[0])
    b3 = gs.shine(withAmplitudes=True)
    scr3 = Screen(bl3, 'scr3', center=(0, 3e4, 20.))
    for nm, lim2 in (('dcm', [-100, 200]), ('dcm_small2', [-100, 123.0])):
        dcm = DCM(bl3, nm, center=(0, 2e4, 0), bragg=theta, material=si, material2=si, cryst2perpTransl=10.,
                  limPhysX=[-50, 50], limPhysY=[-100, 100], limPhysX2=[-0.15, 0.1], limPhysY2=lim2)
        run_both(nm + '_lb', lambda dcm=dcm: dcm.double_reflect(b3, needLocal=True, checkTables=False), dcm, results)
        run_both(nm + '_nolb', lambda dcm=dcm: dcm.double_reflect(b3, needLocal=False, checkTables=False)[:1], dcm,
                 results)
        run_both(nm + '_fused', lambda dcm=dcm: dcm.double_reflect(b3, needLocal=True, checkTables=False,
                                                                   exposeOn=scr3), dcm, results)
    # asymmetric crystal (two normals)
    sia = CrystalSi(hkl=(1, 1, 1))
    dcma = DCM(bl3, 'dcma', center=(0, 2e4, 0), bragg=theta, alpha=np.radians(2.), material=sia, material2=sia,
               cryst2perpTransl=10., limPhysX=[-50, 50], limPhysY=[-100, 100], limPhysX2=[-50, 50],
               limPhysY2=[-100, 200])
    try:
        run_both('dcm_asym_lb', lambda: dcma.double_reflect(b3, needLocal=True, checkTables=False), dcma, results)
    except Exception as e:  # noqa: BLE001
        results['dcm_asym_lb'] = 'skipped: %r' % (e,)
    os.environ.pop('XRTB_REFLECT_TILE', None)
    if HASH_ONLY:
        with open(args.hash_only, 'w') as f:
            json.dump(results, f, indent=0, sort_keys=True)
        return
    ok = all(v == 'ok' or (isinstance(v, str) and v.startswith('skipped')) for v in results.values())
    line = json.dumps({'nrays': n, 'all_ok': ok, 'cases': results})
    print(line)
    os.makedirs('gpurun_out', exist_ok=True)
    with open('gpurun_out/tile_check.jsonl', 'a') as f:
        f.write(line + '\n')
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
