"""f4: spliced surfaces -- the CUDA counterpart of the reference's cl_local_z / cl_local_n text splicing
(oes/base.py:548-569, cl/OE.cl:60-66): BentFlatMirror, ConicalMirror (oes/__init__.py:240-306, :589-636),
JohannCylinder, JohanssonCylinder, JohannToroid (oes/bragg.py:104-269) and a surface written in the test itself.
Oracle bit-exact on the reference's outputs (CPU); plug-in objects build without a GPU (CPU); the fused kernels with
the spliced surface against the same goldens (GPU)."""
import ctypes
import os

import numpy as np
import pytest

from conftest import golden
from helpers import (oracle_spec, oracle_beam, product_beamline, product_oe_kwargs, device_beam, compare_beams,
                     BEAM_FIELDS)
from oracle import materials as om, reflect as orf, surfaces as osf

OUT_FIELDS = BEAM_FIELDS + ('theta', 's', 'phi', 'r')
MIRRORS = ['bentflat', 'conical', 'parab_p', 'parab_q_cyl']
PARAM_KEYS = ('parabParam', 'y0', 'z0', 'cosGamma', 'sinGamma')
CRYSTALS = ['johann_circ', 'johann_parab_alpha', 'johansson', 'johansson_alpha', 'johann_toroid',
            'johann_toroid_alpha']


def _oracle_surface(name, g):
    alpha = (float(g['alpha']) or None) if 'alpha' in g.files else None
    if name == 'bentflat':
        return osf.BentFlatMirror(float(g['R']), float(g['surfPhysY'][0])), alpha
    if name == 'conical':
        return osf.ConicalMirror(float(g['L0']), float(g['theta'])), alpha
    if name.startswith('parab'):
        return osf.ParabolicalMirrorParam(*[float(g[k]) for k in PARAM_KEYS], bool(g['isCylindrical'])), alpha
    if name.startswith('johann_toroid'):
        return osf.JohannToroid(float(g['Rm']), float(g['Rs']), alpha), alpha
    cls = osf.JohanssonCylinder if name.startswith('johansson') else osf.JohannCylinder
    return cls(float(g['Rm']), 'parabolic' if 'parab' in name else 'circular', alpha), alpha


@pytest.mark.parametrize('name', MIRRORS + CRYSTALS)
def test_oracle_spliced_surfaces_bit_exact(name):
    g = golden('reflect_spliced_' + name)
    obj, alpha = _oracle_surface(name, g)
    mat = om.Material('Au', rho=float(g['rho'])) if name in MIRRORS else \
        om.CrystalDiamond((1, 1, 1), float(g['d']), 'Si')
    sp = oracle_spec(g, 'custom', mat, surfParams={'obj': obj}, alpha=alpha, isParametric=name.startswith('parab'))
    info = {}
    gb, lb = orf.reflect(sp, oracle_beam(g, 'in_'), info=info)
    for pre, b in (('gb_', gb), ('lb_', lb)):
        for f in OUT_FIELDS:
            if pre + f in g.files:
                assert np.array_equal(getattr(b, f), g[pre + f], equal_nan=True), pre + f
    assert np.array_equal(info['footprint'], g['footprint'])
    assert (g['gb_state'] == 1).sum() > 300


def _product_oe(name, g):
    from xrt_b200 import oes, materials
    bl = product_beamline(g)
    kw = product_oe_kwargs(g)
    alpha = (float(g['alpha']) or None) if 'alpha' in g.files else None
    if name in MIRRORS:
        kw['material'] = materials.Material('Au', rho=float(g['rho']))
    else:
        kw['material'] = materials.CrystalSi(hkl=(1, 1, 1), tK=float(g['tK']))
        kw['alpha'] = alpha
    if name == 'bentflat':
        oe = oes.BentFlatMirror(bl, name, R=float(g['R']), **kw)
    elif name == 'conical':
        oe = oes.ConicalMirror(bl, name, L0=float(g['L0']), theta=float(g['theta']), **kw)
    elif name.startswith('parab'):
        pq = dict(p=float(g['p'])) if np.isfinite(g['p']) else dict(q=float(g['q']))
        oe = oes.ParabolicalMirrorParam(bl, name, isCylindrical=bool(g['isCylindrical']), **pq, **kw)
        # the host mirror derives the parabola from p | q and the pitch like ParabolicalMirrorParam._reset_pq
        for k in PARAM_KEYS:
            assert abs(getattr(oe, k) - float(g[k])) <= 1e-13 * max(1., abs(float(g[k]))), k
            setattr(oe, k, float(g[k]))
    elif name.startswith('johann_toroid'):
        oe = oes.JohannToroid(bl, name, Rm=float(g['Rm']), Rs=float(g['Rs']), **kw)
    elif name.startswith('johansson'):
        oe = oes.JohanssonCylinder(bl, name, Rm=float(g['Rm']), **kw)
    else:
        oe = oes.JohannCylinder(bl, name, Rm=float(g['Rm']),
                                crossSection='parabolic' if 'parab' in name else 'circular', **kw)
    oe.lostNum = int(g['lostNum'])
    return oe


def test_plugin_objects_build_and_export_without_gpu():
    """nvcc compiles the plug-in of every library surface here (no GPU needed); each object exports the three
    surface-dependent entry points and was built against the current xrt_oe_params."""
    from xrt_b200 import custom, oes_spliced
    from xrt_b200.oes import xrt_oe_params
    paths = oes_spliced.prebuild()
    assert len(set(paths)) == len({(c.cu_local_z, c.cu_local_n) for c in oes_spliced.LIBRARY}) + 5  # + 5 figure-error variants
    assert sum(custom.SurfacePlugin(p).own_intersection for p in paths) == 1  # BlazedGrating
    plugs = [custom.SurfacePlugin(p) for p in paths]
    assert sorted(q.kind for q in plugs).count(5) == 1  # ParabolicalMirrorParam
    assert sorted((q.kind, q.figure_error) for q in plugs if q.figure_error) == [(0, True), (1, True), (2, True),
                                                                                 (3, True), (4, True)]
    for p in paths:
        assert os.path.dirname(p) == custom.PLUGIN_DIR and os.path.exists(p)
        plug = custom.SurfacePlugin(p)
        assert plug.kind in (0, 1, 2, 3, 4, 5)
        assert plug.lib.xrt_plugin_sizeof_oe_params() == ctypes.sizeof(xrt_oe_params)


def test_plugin_compile_error_is_reported():
    from xrt_b200 import custom, _lib
    with pytest.raises(_lib.XrtB200Error) as e:
        custom.build_plugin('__device__ double local_z(const double* p, double x, double y) { return nonsense; }',
                            '__device__ void local_n(const double* p, double x, double y, double* n) {}')
    assert 'nonsense' in str(e.value)


def test_library_has_no_spliced_kinds():
    """libxrt_b200.so itself refuses surface kinds 4 / 5 (they live in plug-in objects) -- checked on the host
    before any launch, so this needs no GPU."""
    from xrt_b200 import _lib
    from xrt_b200.oes import xrt_oe_params
    from xrt_b200.beams import xrt_beam
    lib = _lib.load()
    p = xrt_oe_params()
    p.surface = 4
    b = xrt_beam()
    red = (ctypes.c_uint64 * 64)()
    rc = lib.xrt_reflect_f64(0, ctypes.byref(p), None, None, 0, ctypes.byref(b), None, ctypes.byref(b), None,
                             ctypes.addressof(red), 2, 0, None)
    assert rc != 0 and b'plug-in' in lib.xrt_b200_last_error()


RTOL = 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize('name', MIRRORS + CRYSTALS)
def test_cuda_spliced_reflect_vs_reference_golden(name, cuda_device):
    g = golden('reflect_spliced_' + name)
    oe = _product_oe(name, g)
    gb, lb = oe.reflect(device_beam(oracle_beam(g, 'in_')), withFootprint=True)
    # crystals: amplitudes on the flanks of the Bragg peak carry the 1e4 amplification of the reference's own
    # cancellations (as for the flat crystals): 1e-11
    tol = RTOL if name in MIRRORS else 1e-11
    bad = []
    good = g['lb_state'] == 1
    for pre, b in (('gb_', gb), ('lb_', lb)):
        ref = {f: g[pre + f] for f in OUT_FIELDS if pre + f in g.files}
        dev = b.to_numpy()
        if name.startswith('parab'):
            # parametric surfaces go through atan2 / cos, where CUDA and glibc differ by an ulp: the secant iteration
            # may stop 1-2 ulp of t away, and exp(i k t) amplifies that (same treatment and bounds as for the built-in
            # parametric ellipse, tests/test_reflect.py)
            from xrt_b200.physconsts import CHBAR
            dpath = dev['path'] - ref['path']
            assert np.abs(dpath[good]).max() <= 4e-12
            comp = np.exp(-1e7j * ref['E'] / CHBAR * dpath)
            dev = dict(dev, Es=dev['Es'] * comp, Ep=dev['Ep'] * comp)
            phase_ulp = np.spacing((1e7 * ref['E'] / CHBAR * ref['path']).max())
            bad += compare_beams(dev, ref, ('Es', 'Ep'), 2 * phase_ulp, pre)
            bad += compare_beams(dev, ref, [f for f in OUT_FIELDS if f not in ('Es', 'Ep')], tol, pre)
            continue
        bad += compare_beams(dev, ref, OUT_FIELDS, tol, pre)
    assert not bad, bad
    np.testing.assert_allclose(oe.footprint[0], g['footprint'], rtol=1e-12, atol=1e-12)
    if not name.startswith('parab'):
        assert np.abs(lb.to_numpy()['path'][good] - g['lb_path'][good]).max() <= 2e-12


@pytest.mark.gpu
def test_cuda_user_written_surface_equals_builtin_toroid(cuda_device):
    """A toroid written as cu_local_z / cu_local_n in the test, in the operation order of ToroidMirror
    (oes/__init__.py:398-411), gives the built-in toroid kernel's beams (same solver, same material code)."""
    from xrt_b200 import oes, materials

    class MyToroid(oes.OE):
        cu_plist = ('R', 'r')
        cu_local_z = '''
        __device__ double local_z(const double* cu_plist, double x, double y)
        {
            const double R = cu_plist[0], r = cu_plist[1];
            const double q = x / r;
            double rx = 1.0 - q * q;
            if (rx < 0.0) rx = 0.0;
            return y * y / 2.0 / R + r * (1.0 - sqrt(rx));
        }'''
        cu_local_n = '''
        __device__ void local_n(const double* cu_plist, double x, double y, double* n)
        {
            const double R = cu_plist[0], r = cu_plist[1];
            const double q = x / r;
            const double rx = 1.0 - q * q;
            const double ax = (rx < 0.0) ? 0.0 : 1.0 / sqrt(rx);
            const double a = -x / r * ax, b = -y / R;
            const double norm = sqrt(a * a + b * b + 1.0);
            n[0] = a / norm; n[1] = b / norm; n[2] = 1.0 / norm;
        }'''

        def __init__(self, *a, **k):
            self.R, self.r = k.pop('R'), k.pop('r')
            oes.OE.__init__(self, *a, **k)

    g = golden('reflect_toroid_Au')
    kw = product_oe_kwargs(g)
    mat = materials.Material('Au', rho=float(g['rho']))
    beam = device_beam(oracle_beam(g, 'in_'))
    mine = MyToroid(product_beamline(g), 'mine', R=float(g['R']), r=float(g['r']), material=mat, **kw)
    ref = oes.ToroidMirror(product_beamline(g), 'ref', R=float(g['R']), r=float(g['r']), material=mat, **kw)
    gb1, lb1 = mine.reflect(beam)
    gb2, lb2 = ref.reflect(beam)
    bad = compare_beams(gb1.to_numpy(), gb2.to_numpy(), OUT_FIELDS, RTOL, 'gb_')
    bad += compare_beams(lb1.to_numpy(), lb2.to_numpy(), OUT_FIELDS, RTOL, 'lb_')
    assert not bad, bad
    assert np.array_equal(gb1.to_numpy()['state'], g['gb_state'])
    # the wave-mode helpers go through the same plug-in
    x = beam.x[:64].clone() * 0 + 1.5
    y = beam.y[:64].clone() * 0 + 100.
    assert float((mine.local_z_device(x, y) - ref.local_z_device(x, y)).abs().max()) <= 1e-15
