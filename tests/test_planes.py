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


@pytest.mark.parametrize('name', ['propagate_round', 'propagate_round_stop'])
def test_oracle_propagate_round_bit_exact(name):
    """RoundAperture.propagate (apertures.py:769-825): a pipe and a round beam stop."""
    g = golden(name)
    b = oracle_beam(g, 'in_')
    lo = op.propagate_round(b, list(g['center']), _basis(g), float(g['r']), int(g['lostNum']),
                            isBeamStop=bool(g['isBeamStop']))
    for f in BEAM_FIELDS:
        assert np.array_equal(getattr(lo, f), g['out_' + f]), f
    assert np.array_equal(b.state, g['in_state_after']) and np.array_equal(lo.r, g['lo_r'])
    assert (lo.state == 1).sum() > 500 and (lo.state == int(g['lostNum'])).sum() > 500


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['propagate_round', 'propagate_round_stop'])
def test_cuda_propagate_round_vs_golden(name, cuda_device):
    from xrt_b200.planes import RoundAperture
    g = golden(name)
    ap = RoundAperture(None, 'pipe', center=list(g['center']), r=float(g['r']), isBeamStop=bool(g['isBeamStop']))
    ap.x, ap.y, ap.z = list(g['ex']), list(g['ey']), list(g['ez'])
    ap.lostNum = int(g['lostNum'])
    beam = device_beam(oracle_beam(g, 'in_'))
    lo = ap.propagate(beam)
    assert not compare_beams(lo.to_numpy(), {f: g['out_' + f] for f in BEAM_FIELDS}, BEAM_FIELDS, RTOL)
    assert np.array_equal(beam.state.cpu().numpy(), g['in_state_after'])  # the incoming beam is marked too


@pytest.mark.gpu
def test_cuda_round_prepare_wave_vs_golden(cuda_device):
    """RoundAperture.prepare_wave (apertures.py:848-875) with the reference's uniform deviates."""
    import types
    from xrt_b200.planes import RoundAperture
    g = golden('prepare_wave_round')
    bl = types.SimpleNamespace(sinAzimuth=float(g['sinAzimuth']), cosAzimuth=float(g['cosAzimuth']))
    ap = RoundAperture(None, 'pipe', center=list(g['center']), r=float(g['r']))
    ap.x, ap.y, ap.z = list(g['ex']), list(g['ey']), list(g['ez'])
    src = types.SimpleNamespace(center=list(g['src_center']), bl=bl, name='src')
    wave = ap.prepare_wave(src, 700, xy=g['xy'])
    d = wave.to_numpy()
    for k in ('x', 'z'):
        assert np.abs(d[k] - g[k]).max() <= 1e-15 * float(g['r']) * 4, k  # r cos(phi), r sin(phi): numpy on the host
    for k in ('xDiffr', 'yDiffr', 'zDiffr', 'rDiffr'):
        assert np.abs(d[k] - g[k]).max() <= 1e-12 * 4000., k
    for k in ('a', 'b', 'c'):
        assert np.abs(d[k] - g[k]).max() <= 1e-12, k
    assert wave.area == float(g['area']) and wave.dS == float(g['dS'])


def test_oracle_expose_hemispheric_bit_exact():
    """HemisphericScreen.expose (screens.py:501-545)."""
    g = golden('expose_hemispheric')
    with np.errstate(invalid='ignore'):
        blo = op.expose_hemispheric(oracle_beam(g, 'in_'), list(g['center']), _basis(g), float(g['R']),
                                    int(g['lostNum']), float(g['phiOffset']), float(g['thetaOffset']))
    for f in BEAM_FIELDS + ('theta',):
        assert np.array_equal(getattr(blo, f), g['out_' + f], equal_nan=True), f
    assert np.array_equal(blo.phi, g['out_phi'], equal_nan=True)
    assert (blo.state == int(g['lostNum'])).sum() > 100


@pytest.mark.gpu
def test_cuda_expose_hemispheric_vs_golden(cuda_device):
    import types
    from xrt_b200.planes import HemisphericScreen
    g = golden('expose_hemispheric')
    bl = types.SimpleNamespace(sinAzimuth=float(g['sinAzimuth']), cosAzimuth=float(g['cosAzimuth']), screens=[])
    scr = HemisphericScreen(bl, 'hemi', center=list(g['center']), R=float(g['R']), phiOffset=float(g['phiOffset']),
                            thetaOffset=float(g['thetaOffset']))
    for k, v in (('x', 'ex'), ('y', 'ey'), ('z', 'ez')):  # the default orientation is the reference's
        assert np.allclose(getattr(scr, k), g[v], rtol=0, atol=1e-15), k
    scr.lostNum = int(g['lostNum'])
    blo = scr.expose(device_beam(oracle_beam(g, 'in_'))).to_numpy()
    ref = {f: g['out_' + f] for f in BEAM_FIELDS}
    assert not compare_beams(blo, ref, BEAM_FIELDS, RTOL)
    ok = g['out_state'] > 0
    assert np.abs(blo['theta'][ok] - g['out_theta'][ok]).max() <= 1e-12
    assert np.abs(blo['phi'][ok] - g['out_phi'][ok]).max() <= 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize('n', [0, 1, 1023, 1024, 1025, 300007])
def test_beam_to_host_compacts_selected_rays_in_order(n, cuda_device):
    """Beam.to_host(fields, good_only=True): xrt_compact_rays_f64 == numpy boolean indexing (what raycing.get_output
    does on the host, raycing/__init__.py:170-300), ray order kept, complex fields included."""
    import torch
    from xrt_b200.beams import Beam
    rng = np.random.RandomState(n + 5)
    arrs = dict(x=rng.normal(size=n), z=rng.normal(size=n), E=rng.uniform(8e3, 9e3, n), Jss=rng.rand(n), Jpp=rng.rand(n),
                y=np.zeros(n), a=np.zeros(n), b=np.ones(n), c=np.zeros(n), path=rng.rand(n),
                Jsp=rng.normal(size=n) + 1j * rng.normal(size=n), Es=rng.normal(size=n) + 1j * rng.normal(size=n),
                Ep=rng.normal(size=n) * (1 + 0j), state=rng.choice([1, 1, 1, 2, 3, -5, -2001], n).astype(np.int32))
    b = Beam.from_numpy(arrs)
    for flags, sel in (((1,), arrs['state'] == 1), ((1, 2), (arrs['state'] == 1) | (arrs['state'] == 2)),
                       ((4,), arrs['state'] < 0), ((1, 2, 3, 4), np.ones(n, dtype=bool))):
        got = b.to_host(fields=('x', 'z', 'E', 'Jss', 'Jpp', 'Es', 'Jsp', 'path', 'Ep'), rayFlag=flags)
        for k, v in got.items():
            assert np.array_equal(v, arrs[k][sel]), (k, flags)
    whole = b.to_host(fields=('x', 'Es'), good_only=False)
    assert np.array_equal(whole['x'], arrs['x']) and np.array_equal(whole['Es'], arrs['Es'])
    # the asynchronous pair (to_host_begin / finish: compaction on the compute stream, fields on a copy stream), with
    # two handles in flight and with caller-supplied pinned buffers
    h1 = b.to_host_begin(fields=('x', 'Es', 'Jpp'), rayFlag=(1, 2))
    pinned = {k: torch.empty(max(n, 1), dtype=torch.float64, pin_memory=True) for k in ('z', 'E')}
    h2 = b.to_host_begin(fields=('z', 'E'), rayFlag=(4,), out=pinned)
    got2, got1 = h2.finish(), h1.finish()
    sel12 = (arrs['state'] == 1) | (arrs['state'] == 2)
    for k, v in got1.items():
        assert np.array_equal(v, arrs[k][sel12]), k
    for k, v in got2.items():
        assert np.array_equal(v, arrs[k][arrs['state'] < 0]), k
