"""Full-size (BASELINE.json configs[1]: 1e8 rays) and sharding properties on the GPU: things that must hold at
sizes the numpy oracle cannot reach."""
import numpy as np
import pytest


def _beamline(n, uniform=True):
    import bench
    from xrt_b200.beamline import BeamLine
    from xrt_b200.sources import Undulator
    from xrt_b200.oes import ToroidMirror
    from xrt_b200.materials import Material
    from xrt_b200.planes import Screen
    bl = BeamLine()
    kw = bench.undulator_kwargs(n)
    kw['uniformRayDensity'] = uniform
    src = Undulator(bl, 'u', eMin=bench.E1 * 0.995, eMax=bench.E1 * 1.005, seed=99, **kw)
    mirror = ToroidMirror(bl, 'tm', material=Material('Au', rho=19.3), **bench.MIRROR)
    screen = Screen(bl, 'scr', center=(0, bench.SCREEN_Y, 25000. * np.tan(2 * 3e-3)))
    return src, mirror, screen


@pytest.mark.gpu
def test_full_size_pipeline_invariants(cuda_device):
    """1e8 rays through Undulator -> ToroidMirror(Au) -> Screen: invariants of the physics and bookkeeping."""
    import torch
    n = 100000000
    src, mirror, screen = _beamline(n)
    beam = src.shine()
    assert len(beam) == n and bool((beam.state == 1).all())
    norm = beam.a**2 + beam.b**2 + beam.c**2
    assert float((norm - 1).abs().max()) < 1e-15
    assert float(beam.E.min()) >= src.E_min and float(beam.E.max()) <= src.E_max
    J0 = (beam.Jss + beam.Jpp).clone()
    assert float((J0 - (beam.Es.abs()**2 + beam.Ep.abs()**2)).abs().max()) <= 1e-12 * float(J0.max())
    gb, lb = mirror.reflect(beam, needLocal=True)
    good = gb.state == 1
    ngood = int(good.sum())
    assert 0.2 * n < ngood < n  # the 20 mm x 1 m mirror at 3 mrad intercepts part of the fan
    states = torch.unique(gb.state).tolist()
    assert set(states) <= {1, 2, 3, mirror.lostNum}
    assert bool((lb.state == gb.state).all())
    # energy is untouched, reflectivity <= 1, directions stay normalised, path grows by ~25 m for the good rays
    assert bool((gb.E == beam.E).all())
    J1 = gb.Jss + gb.Jpp
    assert bool((J1[good] <= J0[good] * (1 + 1e-12)).all())
    assert float(((gb.a**2 + gb.b**2 + gb.c**2)[good] - 1).abs().max()) < 1e-14
    dpath = (gb.path - beam.path)[good]
    assert 24000 < float(dpath.min()) and float(dpath.max()) < 26000
    # rays that missed keep the incoming coordinates (reflect.py:131-134)
    miss = ~((gb.state == 1) | (gb.state == 2))
    assert bool((gb.x[miss] == beam.x[miss]).all()) and bool((gb.Jss[miss] == beam.Jss[miss]).all())
    # the local footprint lies on the toroid within zEps
    R, r = mirror.R, mirror.r
    zsurf = lb.y**2 / 2.0 / R + r * (1 - (1 - (lb.x / r)**2).clamp(min=0)**0.5)
    assert float((lb.z - zsurf)[good].abs().max()) <= 2e-12  # zEps + re-evaluation rounding
    del lb, zsurf
    blo = screen.expose(gb)
    assert bool((blo.y == 0).all())
    assert float((blo.path - gb.path)[good].min()) > 24000
    assert bool((blo.Jss == gb.Jss).all())
    amp = (blo.Es.abs()**2 + blo.Ep.abs()**2)[good]
    assert float((amp - (blo.Jss + blo.Jpp)[good]).abs().max()) <= 1e-10 * float(amp.max())


@pytest.mark.gpu
def test_sharded_reflect_equals_whole_beam(cuda_device):
    """Two ray shards with the slot all-reduce between the passes give bit-identical beams to one pass over the
    whole beam (the multi-GPU scheme of xrt_b200/parallel.py, emulated on one GPU)."""
    import torch
    from xrt_b200 import parallel
    from xrt_b200.beams import Beam
    n = 2000001
    src, mirror, screen = _beamline(n)
    beam = src.shine()
    beam.state[::7] = -3
    beam.state[0] = -3  # the first good ray (whose direction sign matters, base.py:1241) is not ray 0
    gb, lb = mirror.reflect(beam)
    h = 777777
    shards = []
    for sl in (slice(0, h), slice(h, n)):
        b = Beam.__new__(Beam)
        for k, v in beam.__dict__.items():
            setattr(b, k, v[sl].contiguous() if isinstance(v, torch.Tensor) else v)
        shards.append(b)

    # emulate the collective: run every phase on both shards, combine their slot buffers, write the result back
    class TwoRankEmulator(object):
        def __init__(self):
            self.bufs = []

        def __call__(self, red, sums=True):
            self.bufs.append(red)

    outs = []
    reds = [torch.zeros(64, dtype=torch.int64, device='cuda') for _ in shards]
    import ctypes
    from xrt_b200 import _lib
    lib = _lib.load()
    params = mirror._make_params(mirror.pitch, mirror.roll + mirror.positionRoll, mirror.yaw, mirror.dx)
    tab, tlen, offs, lens = mirror.material.device_tables(beam.device)
    mat = mirror.material.c_struct(offs, lens)
    gbs = [s.like() for s in shards]
    lbs = [mirror._new_local(s) for s in shards]
    offsets = [0, h]
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def run(phase):
        for s, g, l, r, off in zip(shards, gbs, lbs, reds, offsets):
            cin, cg, cl = s.c_struct(), g.c_struct(), l.c_struct()
            _lib.check(lib.xrt_reflect_f64(len(s), ctypes.byref(params), ctypes.byref(mat), tab.data_ptr(), tlen,
                                           ctypes.byref(cin), None, ctypes.byref(cg), ctypes.byref(cl),
                                           r.data_ptr(), phase, off, stream))

    def allreduce(sums):
        comb = parallel.combine_slots(torch.stack(reds), sums, reds[0].clone())
        for r in reds:
            r.copy_(comb)

    run(1)
    allreduce(True)
    run(4)
    allreduce(False)
    run(2)
    for name, whole, parts in (('gb', gb, gbs), ('lb', lb, lbs)):
        for k, v in whole.tensors().items():
            cat = torch.cat([getattr(p, k) for p in parts])
            if v.is_complex():
                same = bool((torch.view_as_real(cat) == torch.view_as_real(v)).all())
            else:
                if v.is_floating_point():
                    same = bool(((cat == v) | (torch.isnan(cat) & torch.isnan(v))).all())
                else:
                    same = bool((cat == v).all())
            assert same, (name, k)
