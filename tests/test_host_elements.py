"""Host-side logic of the added elements (no GPU): derived parameters equal the reference's (stored in the goldens),
C-struct marshalling of cu_plist, plug-in selection."""
import ctypes

import numpy as np

from conftest import golden
from helpers import product_beamline, product_oe_kwargs


def _kw(g, drop_limits=False):
    kw = product_oe_kwargs(g)
    if drop_limits:
        kw.pop('limPhysX'), kw.pop('limPhysY')
    return kw


def test_normal_fzp_zone_radii_and_limits():
    from xrt_b200 import oes, materials
    g = golden('reflect_fzp_o1')
    fzp = oes.NormalFZP(product_beamline(g), 'fzp', material=materials.Material('Au', rho=19.3, kind='FZP'),
                        f=float(g['f']), E=float(g['E0']), N=int(g['N']), **_kw(g, True))
    assert np.array_equal(fzp.rn, g['rn'])
    assert fzp.limPhysX == list(g['surfPhysX']) and fzp.limPhysY == list(g['surfPhysY'])
    thin = oes.NormalFZP(None, 'fzp2', material=materials.Material('Au', rho=19.3, kind='FZP'), f=100., E=9000.,
                         thinnestZone=5e-5)
    assert abs(thin.N - (12398.419297617678 / 9000. * 1e-7) * 100. / 4. / 5e-5**2) < 1e-9  # gratings.py:99-100


def test_parabolic_mirror_reset_pq_equals_reference():
    from xrt_b200 import oes
    for name in ('parab_p', 'parab_q_cyl'):
        g = golden('reflect_spliced_' + name)
        pq = dict(p=float(g['p'])) if np.isfinite(g['p']) else dict(q=float(g['q']))
        oe = oes.ParabolicalMirrorParam(product_beamline(g), name, isCylindrical=bool(g['isCylindrical']), **pq,
                                        **_kw(g))
        for k in ('parabParam', 'y0', 'z0', 'cosGamma', 'sinGamma'):
            assert abs(getattr(oe, k) - float(g[k])) <= 1e-13 * max(1., abs(float(g[k]))), (name, k)
        p = oe._make_params(oe.pitch, oe.roll + oe.positionRoll, oe.yaw)
        assert p.surface == 5 and abs(p.sp[1] - oe.parabParam**2) == 0. and p.sp[6] == float(oe.isCylindrical)


def test_cu_plist_marshalling_and_two_normals():
    from xrt_b200 import oes
    from xrt_b200.materials import CrystalSi
    si = CrystalSi(hkl=(1, 1, 1))
    jc = oes.JohannCylinder(None, 'jc', Rm=1500., material=si, alpha=0.1)
    p = jc._make_params(0.2, 0., 0.)
    assert p.surface == 4 and list(p.sp)[:5] == [0., 1500., 1500.**2, float(np.cos(0.1)), float(np.sin(0.1))]
    assert p.asymmetric == 1
    assert oes.JohannCylinder(None, 'jc0', Rm=1500., material=si)._make_params(0.2, 0., 0.).asymmetric == 0
    assert oes.JohanssonCylinder(None, 'js', Rm=1500., material=si)._make_params(0.2, 0., 0.).asymmetric == 1
    lp = oes.LauePlate(None, 'lp', material=CrystalSi(hkl=(1, 1, 1), geom='Laue reflected', t=0.1))
    q = lp._make_params(1.3, 0., 0.)
    assert q.surface == 0 and q.asymmetric == 1 and list(q.nB) == [0., 1., 0.] and lp._plugin() is None
    lpa = oes.LauePlate(None, 'lpa', material=CrystalSi(hkl=(1, 1, 1), geom='Laue reflected', t=0.1), alpha=0.2)
    assert np.allclose(list(lpa._make_params(1.3, 0., 0.).nB), [0., np.cos(0.2), -np.sin(0.2)], rtol=0, atol=1e-16)
    bg = oes.BlazedGrating(None, 'bg', blaze=0.02, rho=300.)
    assert bg.cu_find_intersection and list(bg._make_params(0.03, 0., 0.).sp)[0] == 1. / 300.
    rd = oes.OE(None, 'rd', shape='round', limPhysX=[-2, 4], limPhysY=[-30, 10])._make_params(0.01, 0., 0.)
    assert rd.shape_round == 1 and list(rd.shape_r) == [1., 3., -10., 20.]


def test_plugin_selection_without_gpu():
    """Which object serves which element: built-in surfaces -> the library; spliced texts, figure errors and own
    intersections -> their plug-in variants (built by nvcc here, no GPU needed)."""
    from xrt_b200 import oes, figure_error as pfe
    lims = dict(limPhysX=[-4, 4], limPhysY=[-100, 100])
    assert oes.ToroidMirror(None, 'tm', R=1e6, r=50.)._plugin() is None
    v = oes.BentFlatMirror(None, 'vcm', R=1e6, **lims)._plugin()
    assert v.kind == 4 and not v.figure_error and not v.own_intersection
    fe = pfe.Waviness(amplitude=5., xWaveLength=2., yWaveLength=40., gridStep=8., **lims)
    t = oes.ToroidMirror(None, 'tmfe', R=1e6, r=50., figureError=fe, **lims)._plugin()
    assert t.kind == 1 and t.figure_error
    assert oes.BlazedGrating(None, 'bg', blaze=0.02, rho=300.)._plugin().own_intersection
    assert t.lib.xrt_plugin_sizeof_oe_params() == ctypes.sizeof(oes.xrt_oe_params)
