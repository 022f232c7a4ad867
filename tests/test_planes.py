"""a15: Screen.expose and RectangularAperture.propagate -- oracle bit-exact on reference outputs (CPU), CUDA
vs golden (GPU)."""
import numpy as np
import pytest

from conftest import golden
from helpers import oracle_beam, device_beam, compare_beams, BEAM_FIELDS
from oracle import planes as op

RTOL = 1e-12


def _basis(g):
    return [list(g['ex']), list(g['ey']), list(g['ez'])]


@pytest.mark.parametrize('name', ['expose_screen', 'expose_screen_tilted'])
def test_oracle_expose_bit_exact(name):
    g = golden(name)
    info = {}
    blo = op.expose(oracle_beam(g, 'in_'), list(g['center']), _basis(g), int(g['lostNum']),
                    onlyPositivePath=name.endswith('tilted'),
                    compressX=float(g['compressX']) if 'compressX' in g.files else None, info=info)
    for f in BEAM_FIELDS:
        assert np.array_equal(getattr(blo, f), g['out_' + f]), f
    if 'footprint' in g.files:
        assert np.array_equal(info['footprint'], g['footprint'])


def test_oracle_propagate_bit_exact():
    g = golden('propagate_slit')
    b = oracle_beam(g, 'in_')
    info = {}
    lo = op.propagate_rect(b, list(g['center']), _basis(g), dict(zip([str(k) for k in g['kinds']], g['opening'])),
                           int(g['lostNum']), info=info)
    for f in BEAM_FIELDS:
        assert np.array_equal(getattr(lo, f), g['out_' + f]), f
    assert np.array_equal(b.state, g['in_state_after'])
    assert np.allclose(info['spotLimits'], g['spotLimits'], rtol=0, atol=0)


def _screen(g):
    from xrt_b200.planes import Screen
    from helpers import product_beamline
    scr = Screen(None, 'scr', center=list(g['center']), x=list(g['ex']), z=list(g['ez']),
                 compressX=float(g['compressX']) if 'compressX' in g.files else None)
    scr.x, scr.y, scr.z = list(g['ex']), list(g['ey']), list(g['ez'])  # exactly the reference's axes
    scr.lostNum = int(g['lostNum'])
    return scr


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['expose_screen', 'expose_screen_tilted'])
def test_cuda_expose_vs_golden(name, cuda_device):
    g = golden(name)
    scr = _screen(g)
    blo = scr.expose(device_beam(oracle_beam(g, 'in_')), onlyPositivePath=name.endswith('tilted'),
                     withFootprint=True)
    ref = {f: g['out_' + f] for f in BEAM_FIELDS}
    assert not compare_beams(blo.to_numpy(), ref, BEAM_FIELDS, RTOL)
    if 'footprint' in g.files:
        np.testing.assert_allclose(scr.footprint[0], g['footprint'], rtol=1e-13, atol=1e-13)


@pytest.mark.gpu
def test_cuda_screen_auto_axes_match_reference(cuda_device):
    """Screen(bl, x='auto', z='auto') derives its axes from the beamline azimuth (beamline.py:290-318)."""
    from xrt_b200.planes import Screen
    from xrt_b200.beamline import BeamLine
    g = golden('expose_screen')
    bl = BeamLine()
    bl.sinAzimuth, bl.cosAzimuth = float(g['sinAzimuth']), float(g['cosAzimuth'])
    scr = Screen(bl, 's', center=list(g['center']))
    assert np.array_equal(np.array(scr.x, dtype=float), g['ex'])
    assert np.array_equal(np.array(scr.y, dtype=float), g['ey'])
    assert np.array_equal(np.array(scr.z, dtype=float), g['ez'])
    assert scr.lostNum == int(g['lostNum'])


@pytest.mark.gpu
def test_cuda_propagate_vs_golden(cuda_device):
    from xrt_b200.planes import RectangularAperture
    g = golden('propagate_slit')
    slit = RectangularAperture(None, 'slit', center=list(g['center']), kind=[str(k) for k in g['kinds']],
                               opening=list(g['opening']))
    slit.x, slit.y, slit.z = list(g['ex']), list(g['ey']), list(g['ez'])
    slit.lostNum = int(g['lostNum'])
    beam = device_beam(oracle_beam(g, 'in_'))
    lo = slit.propagate(beam, withSpotLimits=True)
    ref = {f: g['out_' + f] for f in BEAM_FIELDS}
    assert not compare_beams(lo.to_numpy(), ref, BEAM_FIELDS, RTOL)
    assert np.array_equal(beam.state.cpu().numpy(), g['in_state_after'])  # the incoming beam is marked too
    np.testing.assert_allclose(slit.spotLimits, g['spotLimits'], rtol=1e-13, atol=1e-13)
