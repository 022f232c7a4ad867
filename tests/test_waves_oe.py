"""a16/a18 with an OE as the receiving surface: OE.prepare_wave (oes/reflect.py:266-403), diffract onto the OE
(waves.py:786-824), reflect(noIntersectionSearch=True) (reflect.py:676-681) and on to a screen -- the body of
OE.propagate_wave (:405-449).  Golden: the unmodified reference on slit -> ToroidMirror (random samples) and
slit -> EllipticalMirrorParam (mesh samples, some of them lost) -> Screen (tools/gen_golden.py wave_oe)."""
import numpy as np
import pytest

from conftest import golden
from helpers import oracle_beam, device_beam, compare_beams, oracle_spec, product_beamline, product_oe_kwargs
from oracle import waves as ow, reflect as orf, materials as om

WAVE_FIELDS = ('x', 'y', 'z', 'a', 'b', 'c', 'E', 'Jss', 'Jpp', 'Jsp', 'Es', 'Ep', 'state')
GLO_FIELDS = ('x', 'y', 'z', 'a', 'b', 'c', 'E', 'Jss', 'Jpp', 'Jsp', 'Es', 'Ep', 'state')
REFL_FIELDS = ('x', 'y', 'z', 'a', 'b', 'c', 'path', 'E', 'Jss', 'Jpp', 'Jsp', 'Es', 'Ep', 'state')
ACC = ('EsAcc', 'EpAcc', 'aEacc', 'bEacc', 'cEacc')
CASES = ('wave_slit_toroid_screen', 'wave_slit_kb_screen')


def _samples(g):
    px, py = g['surfPhysX'], g['surfPhysY']
    if 'xy' in g.files:
        xy = g['xy']
        return xy[:, 0] * (px[1] - px[0]) + px[0], xy[:, 1] * (py[1] - py[0]) + py[0]
    nx, ny = [int(v) for v in g['mesh']]
    X, Y = np.meshgrid(np.linspace(px[0], px[1], nx), np.linspace(py[0], py[1], ny))
    return X.ravel(), Y.ravel()


def _oracle_mirror(g):
    Au = om.Material('Au', rho=float(g['rho']))
    if 'ellipseA' in g.files:
        return oracle_spec(g, 'ellipse_param', Au, isParametric=True,
                           surfParams=dict(isCylindrical=True, **{k: float(g[k]) for k in
                                           ('ellipseA', 'ellipseB', 'y0', 'z0', 'cosGamma', 'sinGamma')}))
    return oracle_spec(g, 'toroid', Au, surfParams=dict(R=float(g['R']), r=float(g['r'])))


@pytest.mark.parametrize('name', CASES)
def test_oracle_wave_onto_oe_bit_exact(name):
    g = golden(name)
    mirror = _oracle_mirror(g)
    slit = ow.Plane(list(g['slit_center']), float(g['sinAzimuth']), float(g['cosAzimuth']))
    x, y = _samples(g)
    wave = ow.oe_prepare_wave(mirror, slit, x, y)
    for f in ('x', 'y', 'z', 'a', 'b', 'c', 'state', 'xDiffr', 'yDiffr', 'zDiffr', 'rDiffr'):
        assert np.array_equal(getattr(wave, f), g['pw_' + f]), 'pw_' + f
    assert wave.area == float(g['pw_area']) and wave.dS == float(g['pw_dS'])
    assert wave.areaNormal == float(g['pw_areaNormal'])
    oeLocal = oracle_beam(g, 'in_')
    glo = ow.diffract(oeLocal, wave)
    ow.to_oe_frames(wave, glo)
    assert oeLocal.area == float(g['slit_area'])
    for f in ACC:
        assert np.array_equal(getattr(wave, f), g['wave_' + f]), f
    for f in WAVE_FIELDS:
        assert np.array_equal(getattr(wave, f), g['wave_' + f]), 'wave_' + f
    for f in GLO_FIELDS:
        assert np.array_equal(getattr(glo, f), g['glo_' + f]), 'glo_' + f
    mglo, mloc = orf.reflect(mirror, glo, createdByDiffract=True, noIntersectionSearch=True)
    for f in REFL_FIELDS:
        assert np.array_equal(getattr(mglo, f), g['mglo_' + f]), 'mglo_' + f
        assert np.array_equal(getattr(mloc, f), g['mloc_' + f]), 'mloc_' + f
    scr = ow.Plane(list(g['scr_center']), float(g['sinAzimuth']), float(g['cosAzimuth']))
    wave2 = ow.screen_prepare_wave_from_oe(scr, mirror, g['xmesh'], g['zmesh'])
    glo2 = ow.diffract_from_oe(mloc, wave2)
    assert mloc.area == float(g['mirror_area'])
    for f in ACC:
        assert np.array_equal(getattr(wave2, f), g['wave2_' + f]), f
    for f in GLO_FIELDS:
        assert np.array_equal(getattr(glo2, f), g['glo2_' + f]), 'glo2_' + f


def _product_mirror(g, bl):
    from xrt_b200 import oes, materials
    Au = materials.Material('Au', rho=float(g['rho']))
    if 'ellipseA' in g.files:
        kb = oes.EllipticalMirrorParam(bl, 'kb', p=float(g['p']), q=float(g['q']), isCylindrical=True, material=Au,
                                       **product_oe_kwargs(g))
        for k in ('ellipseA', 'ellipseB', 'y0', 'z0', 'cosGamma', 'sinGamma'):
            assert abs(getattr(kb, k) - float(g[k])) <= 1e-13 * max(1., abs(float(g[k]))), k
            setattr(kb, k, float(g[k]))
        return kb
    return oes.ToroidMirror(bl, 'tm', R=float(g['R']), r=float(g['r']), material=Au, **product_oe_kwargs(g))


@pytest.mark.gpu
@pytest.mark.parametrize('name', CASES)
def test_cuda_wave_onto_oe_vs_golden(cuda_device, name):
    from xrt_b200 import waves
    from xrt_b200.planes import Screen, RectangularAperture
    g = golden(name)
    bl = product_beamline(g)
    mirror = _product_mirror(g, bl)
    slit = RectangularAperture(bl, 'slit', center=list(g['slit_center']), kind=('left', 'right', 'bottom', 'top'),
                               opening=list(g['opening']))
    if 'xy' in g.files:
        wave = mirror.prepare_wave(slit, int(len(g['xy'])), xy=g['xy'])
    else:
        wave = mirror.prepare_wave(slit, tuple(int(v) for v in g['mesh']))
    w = wave.to_numpy()
    assert np.array_equal(w['state'], g['pw_state'])
    # sample positions on the surface: the solver lands within 2e-12 mm of the prepared points
    for f in ('x', 'y', 'z'):
        assert np.abs(w[f] - g['pw_' + f]).max() <= 4e-12, f
    for f in ('xDiffr', 'yDiffr', 'zDiffr', 'rDiffr'):
        np.testing.assert_allclose(w[f], g['pw_' + f], rtol=1e-13, atol=4e-12)
    assert abs(wave.area - float(g['pw_area'])) <= 1e-13 * float(g['pw_area'])
    assert abs(wave.areaNormal - float(g['pw_areaNormal'])) <= 1e-12 * float(g['pw_areaNormal'])
    assert abs(wave.dS - float(g['pw_dS'])) <= 1e-13 * float(g['pw_dS'])
    # the Kirchhoff phase k*r (k = 4.6e7 /mm at 9 keV) turns a 2e-12 mm difference of a receiving point into
    # 1e-4 rad: parity of the field needs the reference's receiving points, exactly like an xrt user who
    # propagates onto the same prepared wave
    for f in ('xDiffr', 'yDiffr', 'zDiffr', 'rDiffr', 'x', 'y', 'z', 'a', 'b', 'c'):
        import torch
        setattr(wave, f, torch.from_numpy(np.ascontiguousarray(g['pw_' + f])).to(wave.device))
    oeLocal = device_beam(oracle_beam(g, 'in_'))
    glo = waves.diffract(oeLocal, wave)
    assert abs(oeLocal.area - float(g['slit_area'])) <= 1e-13 * float(g['slit_area'])
    w = wave.to_numpy()
    for f in ACC:
        ref = g['wave_' + f]
        assert np.abs(w[f] - ref).max() <= 1e-11 * np.abs(ref).max(), f
    bad = compare_beams(w, {f: g['wave_' + f] for f in WAVE_FIELDS}, WAVE_FIELDS, 1e-10, 'wave_')
    bad += compare_beams(glo.to_numpy(), {f: g['glo_' + f] for f in GLO_FIELDS}, GLO_FIELDS, 1e-10, 'glo_')
    assert not bad, bad
    mglo, mloc = mirror.reflect(glo, noIntersectionSearch=True)
    bad = compare_beams(mglo.to_numpy(), {f: g['mglo_' + f] for f in REFL_FIELDS}, REFL_FIELDS, 1e-10, 'mglo_')
    bad += compare_beams(mloc.to_numpy(), {f: g['mloc_' + f] for f in REFL_FIELDS}, REFL_FIELDS, 1e-10, 'mloc_')
    assert not bad, bad
    # mirror -> screen from the reference's local beam (same reason as above)
    scr = Screen(bl, 'scr', center=list(g['scr_center']))
    wave2 = waves.screen_prepare_wave(scr, mirror, g['xmesh'], g['zmesh'])
    mloc_ref = device_beam(oracle_beam(g, 'mloc_'))
    glo2 = waves.diffract(mloc_ref, wave2)
    assert abs(mloc_ref.area - float(g['mirror_area'])) <= 1e-12 * float(g['mirror_area'])
    w2 = wave2.to_numpy()
    for f in ACC:
        ref = g['wave2_' + f]
        assert np.abs(w2[f] - ref).max() <= 1e-11 * np.abs(ref).max(), f
    bad = compare_beams(glo2.to_numpy(), {f: g['glo2_' + f] for f in GLO_FIELDS}, GLO_FIELDS, 1e-10, 'glo2_')
    assert not bad, bad


@pytest.mark.gpu
def test_cuda_propagate_wave_runs_end_to_end(cuda_device):
    """OE.propagate_wave with its own prepared samples (no golden substitution): the intensity on the mirror
    agrees with the reference's run of the same chain within the sampling-independent 1e-6 (phase 1e-4 rad)."""
    from xrt_b200.planes import RectangularAperture
    g = golden('wave_slit_toroid_screen')
    bl = product_beamline(g)
    mirror = _product_mirror(g, bl)
    slit = RectangularAperture(bl, 'slit', center=list(g['slit_center']), kind=('left', 'right', 'bottom', 'top'),
                               opening=list(g['opening']))
    oeLocal = device_beam(oracle_beam(g, 'in_'))
    mglo, mloc = mirror.propagate_wave(oeLocal, prevOE=slit, nrays=int(len(g['xy'])), xy=g['xy'])
    m = mloc.to_numpy()
    assert np.array_equal(m['state'], g['mloc_state'])
    I, Iref = m['Jss'] + m['Jpp'], g['mloc_Jss'] + g['mloc_Jpp']
    assert np.abs(I - Iref).max() <= 2e-3 * Iref.max()
    assert abs(I.sum() - Iref.sum()) <= 1e-4 * Iref.sum()
