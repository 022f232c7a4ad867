"""a16-a18: Screen.prepare_wave + waves.diffract (slit -> screen) -- oracle bit-exact on the reference's
rw.diffract outputs (CPU), product (CUDA Kirchhoff kernel + device bookkeeping) against the same golden (GPU)."""
import numpy as np
import pytest

from conftest import golden
from helpers import oracle_beam, device_beam, compare_beams
from oracle import waves as ow

WAVE_FIELDS = ('a', 'b', 'c', 'E', 'Jss', 'Jpp', 'Jsp', 'Es', 'Ep', 'state')
GLO_FIELDS = ('x', 'y', 'z', 'a', 'b', 'c', 'E', 'Jss', 'Jpp', 'Jsp', 'Es', 'Ep', 'state')


def _planes(g):
    sa, ca = np.sin(float(g['azimuth'])), np.cos(float(g['azimuth']))
    return ow.Plane(list(g['slit_center']), sa, ca), ow.Plane(list(g['scr_center']), sa, ca)


def test_oracle_prepare_wave_and_diffract_bit_exact():
    g = golden('diffract_slit_screen')
    slit, scr = _planes(g)
    wave = ow.screen_prepare_wave(scr, slit, g['xmesh'], g['zmesh'])
    for k, gk in (('xDiffr', 'xDiffr'), ('yDiffr', 'yDiffr'), ('zDiffr', 'zDiffr'), ('rDiffr', 'rDiffr'), ('a', 'wa'),
                  ('b', 'wb'), ('c', 'wc')):
        assert np.array_equal(getattr(wave, k), g[gk]), k
    assert wave.dS == float(g['dS']) and wave.area == float(g['warea'])
    oeLocal = oracle_beam(g, 'in_')
    glo = ow.diffract(oeLocal, wave)
    assert oeLocal.area == float(g['area'])
    for f in WAVE_FIELDS:
        assert np.array_equal(getattr(wave, f), g['wave_' + f]), 'wave_' + f
    for f in ('EsAcc', 'EpAcc', 'aEacc', 'bEacc', 'cEacc'):
        assert np.array_equal(getattr(wave, f), g['wave_' + f]), f
    for f in GLO_FIELDS:
        assert np.array_equal(getattr(glo, f), g['glo_' + f]), 'glo_' + f


@pytest.mark.gpu
def test_cuda_prepare_wave_and_diffract_vs_golden(cuda_device):
    from xrt_b200 import waves
    from xrt_b200.beamline import BeamLine
    from xrt_b200.planes import Screen, RectangularAperture
    g = golden('diffract_slit_screen')
    bl = BeamLine(azimuth=float(g['azimuth']))
    slit = RectangularAperture(bl, 'slit', center=list(g['slit_center']), kind=('left', 'right', 'bottom', 'top'),
                               opening=list(g['opening']))
    scr = Screen(bl, 'scr', center=list(g['scr_center']))
    wave = waves.screen_prepare_wave(scr, slit, g['xmesh'], g['zmesh'])
    for k, gk in (('xDiffr', 'xDiffr'), ('yDiffr', 'yDiffr'), ('zDiffr', 'zDiffr'), ('rDiffr', 'rDiffr'), ('a', 'wa'),
                  ('b', 'wb'), ('c', 'wc')):
        np.testing.assert_allclose(getattr(wave, k).cpu().numpy(), g[gk], rtol=1e-13, atol=1e-13)
    assert abs(wave.dS - float(g['dS'])) < 1e-18 and abs(wave.area - float(g['warea'])) < 1e-12
    oeLocal = device_beam(oracle_beam(g, 'in_'))
    glo = waves.diffract(oeLocal, wave)
    assert abs(oeLocal.area - float(g['area'])) < 1e-15
    w = wave.to_numpy()
    # accumulators: 1e-12 of the largest amplitude (the sum itself is checked against sum|terms| in
    # test_kirchhoff.py); derived fields: 1e-11 of their group scale (angle()/exp() of the de-phasing)
    for f in ('EsAcc', 'EpAcc', 'aEacc', 'bEacc', 'cEacc'):
        ref = g['wave_' + f]
        assert np.abs(w[f] - ref).max() <= 1e-12 * np.abs(ref).max(), f
    bad = compare_beams(w, {f: g['wave_' + f] for f in WAVE_FIELDS}, WAVE_FIELDS, 1e-11, 'wave_')
    bad += compare_beams(glo.to_numpy(), {f: g['glo_' + f] for f in GLO_FIELDS}, GLO_FIELDS, 1e-11, 'glo_')
    assert not bad, bad
    assert abs(wave.beamReflSumJ - float(g['beamReflSumJ'])) <= 1e-12 * float(g['beamReflSumJ'])


def _kb_oracle(g):
    from helpers import oracle_spec
    from oracle import materials as om
    return oracle_spec(g, 'ellipse_param', om.Material('Au', rho=19.3), isParametric=True,
                       surfParams=dict(isCylindrical=True, **{k: float(g[k]) for k in
                                       ('ellipseA', 'ellipseB', 'y0', 'z0', 'cosGamma', 'sinGamma')}))


def test_oracle_diffract_from_kb_mirror_bit_exact():
    """Config 5, last step: the rays traced onto an elliptical KB mirror are the diffracting samples."""
    g = golden('diffract_kb_screen')
    kb = _kb_oracle(g)
    scr = ow.Plane(list(g['scr_center']), float(g['sinAzimuth']), float(g['cosAzimuth']))
    wave = ow.screen_prepare_wave_from_oe(scr, kb, g['xmesh'], g['zmesh'])
    for k in ('xDiffr', 'yDiffr', 'zDiffr'):
        assert np.array_equal(getattr(wave, k), g[k]), k
    lb = oracle_beam(g, 'in_')
    glo = ow.diffract_from_oe(lb, wave)
    assert lb.area == float(g['area'])
    for f in WAVE_FIELDS:
        assert np.array_equal(getattr(wave, f), g['wave_' + f]), 'wave_' + f
    for f in ('EsAcc', 'EpAcc', 'aEacc', 'bEacc', 'cEacc'):
        assert np.array_equal(getattr(wave, f), g['wave_' + f]), f
    for f in GLO_FIELDS:
        assert np.array_equal(getattr(glo, f), g['glo_' + f]), 'glo_' + f


@pytest.mark.gpu
def test_cuda_diffract_from_kb_mirror_vs_golden(cuda_device):
    from xrt_b200 import waves, oes, materials
    from xrt_b200.planes import Screen
    from helpers import product_beamline, product_oe_kwargs
    g = golden('diffract_kb_screen')
    bl = product_beamline(g)
    kb = oes.EllipticalMirrorParam(bl, 'kb', p=float(g['p']), q=float(g['q']), isCylindrical=True,
                                   material=materials.Material('Au', rho=19.3), **product_oe_kwargs(g))
    for k in ('ellipseA', 'ellipseB', 'y0', 'z0', 'cosGamma', 'sinGamma'):
        assert abs(getattr(kb, k) - float(g[k])) <= 1e-13 * max(1., abs(float(g[k]))), k
        setattr(kb, k, float(g[k]))
    scr = Screen(bl, 'scr', center=list(g['scr_center']))
    wave = waves.screen_prepare_wave(scr, kb, g['xmesh'], g['zmesh'])
    for k in ('xDiffr', 'yDiffr', 'zDiffr'):
        np.testing.assert_allclose(getattr(wave, k).cpu().numpy(), g[k], rtol=1e-13, atol=1e-11)
    lb = device_beam(oracle_beam(g, 'in_'))
    glo = waves.diffract(lb, wave)
    assert abs(lb.area - float(g['area'])) <= 1e-12 * float(g['area'])
    w = wave.to_numpy()
    for f in ('EsAcc', 'EpAcc', 'aEacc', 'bEacc', 'cEacc'):
        ref = g['wave_' + f]
        assert np.abs(w[f] - ref).max() <= 1e-11 * np.abs(ref).max(), f
    bad = compare_beams(w, {f: g['wave_' + f] for f in WAVE_FIELDS}, WAVE_FIELDS, 1e-10, 'wave_')
    bad += compare_beams(glo.to_numpy(), {f: g['glo_' + f] for f in GLO_FIELDS}, GLO_FIELDS, 1e-10, 'glo_')
    assert not bad, bad
