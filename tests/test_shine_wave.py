"""f3: IntegratedSource.shine(wave=...) (sources/sybase.py:1513-1812) -- an undulator illuminating the receiving
points of a prepared wave, as a filament beam (one electron, energy spread) or with the full emittance.  Oracle
bit-exact on the reference's outputs with the seeded legacy np.random stream (CPU); the CUDA path (device
build_I_map + device elementwise) fed with the oracle's deviates (GPU)."""
import numpy as np
import pytest

from conftest import golden
from helpers import BEAM_FIELDS
from oracle import shine as osh, undulator as ou

CASES = ['shine_wave_filament', 'shine_wave_full']
KEYS = ('L0', 'Np', 'Kx', 'Ky', 'eI', 'gamma', 'phase', 'E_min', 'E_max', 'Theta_min', 'Theta_max', 'Psi_min',
        'Psi_max', 'dx', 'dz', 'dxprime', 'dzprime', 'E1', 'eEspread')


def _spec(g):
    kw = {k: float(g[k]) for k in KEYS}
    return osh.UndulatorSpec(quadm=int(g['quadm']), gIntervals=int(g['gIntervals']), uniformRayDensity=True,
                             center=list(g['center']), azimuth=(float(g['sinAzimuth']), float(g['cosAzimuth'])), **kw)


def _oracle_run(name, record=None):
    g = golden(name)
    src = _spec(g)
    src.Imax = float(g['Imax0'])  # a filament source enters shine with the Imax its reset() probed (sybase.py:529-541)
    src.fluxConst = src.Imax * src.xzE  # ... and the flux constant reset() derived from it (sybase.py:549)
    wave = osh.Wave(g['xDiffr'], g['yDiffr'], g['zDiffr'], float(g['area']))
    wave.x[:], wave.y[:], wave.z[:] = g['wave_in_x'], g['wave_in_y'], g['wave_in_z']  # screen-local, untouched by shine
    np.random.seed(int(g['seed']))
    bor = osh.shine_wave(src, wave, filamentBeam=bool(g['filament']), record=record)
    return g, src, wave, bor


@pytest.mark.parametrize('name', CASES)
def test_oracle_shine_wave_bit_exact(name):
    g, src, wave, bor = _oracle_run(name)
    for pre, b in (('bor_', bor), ('wave_', wave)):
        for f in BEAM_FIELDS:
            assert np.array_equal(getattr(b, f), g[pre + f]), pre + f
    assert np.array_equal(wave.rDiffr, g['rDiffr'])
    for k in ('accepted', 'acceptedE', 'seeded', 'seededI', 'sourceWeight'):
        assert getattr(bor, k) == g[k], k
    assert src.Imax == float(g['Imax']) and src.xzE == float(g['xzE'])
    if bool(g['filament']):
        for k in ('filamentDX', 'filamentDZ', 'filamentDtheta', 'filamentDpsi', 'filamentDgamma'):
            assert getattr(wave, k) == float(g[k]), k


def _product_source(g):
    from xrt_b200.beamline import BeamLine
    from xrt_b200.sources import Undulator
    bl = BeamLine()
    bl.sinAzimuth, bl.cosAzimuth = float(g['sinAzimuth']), float(g['cosAzimuth'])
    E1 = float(g['E1'])
    src = Undulator(bl, 'u', center=list(g['center']), nrays=1000, eE=3., eI=float(g['eI']), period=float(g['L0']),
                    n=int(g['Np']), K=float(g['Ky']), eSigmaX=float(g['dx'])*1e3, eSigmaZ=float(g['dz'])*1e3,
                    eEpsilonX=float(g['dx'])*float(g['dxprime'])*1e6, eEpsilonZ=float(g['dz'])*float(g['dzprime'])*1e6,
                    xPrimeMax=0.1, zPrimeMax=0.1, xPrimeMaxAutoReduce=False, zPrimeMaxAutoReduce=False,
                    eMin=float(g['E_min']), eMax=float(g['E_max']), gNodes=int(g['quadm']),
                    gIntervals=int(g['gIntervals']), eEspread=float(g['eEspread']), uniformRayDensity=True,
                    filamentBeam=bool(g['filament']))
    src.reset()
    for k in ('dx', 'dz', 'dxprime', 'dzprime', 'gamma', 'E1'):  # the golden's exact scalars
        assert abs(getattr(src, k) - float(g[k])) <= 1e-12 * abs(float(g[k])), k
        setattr(src, k, float(g[k]))
    return src, bl, E1


@pytest.mark.gpu
@pytest.mark.parametrize('name', CASES)
def test_cuda_shine_wave_vs_reference_golden(name, cuda_device):
    """Screen.prepare_wave(source, mesh) -> source.shine(wave=wave) on the device with the oracle's deviates: geometry
    1e-12 of the group scales, amplitudes 1e-10 of the largest one (as for shine: the fused prefactor `ab` near its
    zeros); the propagation phase k r ~ 1e11 rad is formed from bit-identical r."""
    import torch
    from xrt_b200.planes import Screen
    rec = {}
    g, osrc, owave, obor = _oracle_run(name, record=rec)
    src, bl, E1 = _product_source(g)
    src.Imax = float(g['Imax0'])
    src.fluxConst = src.Imax * src.xzE
    scr = Screen(bl, 'scr', center=list(g['screen_center']))
    wave = scr.prepare_wave(src, g['xm'], g['zm'])
    for k in ('xDiffr', 'yDiffr', 'zDiffr'):
        assert np.abs(getattr(wave, k).cpu().numpy() - g[k]).max() <= 1e-12 * 25000.
    assert abs(wave.area - float(g['area'])) <= 1e-12 * float(g['area'])
    dev = wave.device
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    if bool(g['filament']):
        samples = {k: float(v) for k, v in rec.items() if k in ('uE', 'gX', 'gZ', 'gT', 'gP', 'gG')}
    else:
        samples = {k: t(rec[k]) for k in ('uE', 'gX', 'gZ', 'gT', 'gP', 'g1', 'g2')}
    bor = src.shine(wave=wave, samples=samples)
    from helpers import compare_beams
    dw, db = wave.to_numpy(), bor.to_numpy()
    geo = [f for f in BEAM_FIELDS if f not in ('Es', 'Ep', 'Jss', 'Jpp', 'Jsp')]
    bad = compare_beams(dw, {f: g['wave_' + f] for f in BEAM_FIELDS}, geo, 1e-12, 'wave_')
    bad += compare_beams(db, {f: g['bor_' + f] for f in BEAM_FIELDS}, geo, 1e-12, 'bor_')
    bad += compare_beams(dw, {f: g['wave_' + f] for f in BEAM_FIELDS}, ('Es', 'Ep', 'Jss', 'Jpp', 'Jsp'), 1e-10, 'wave_')
    bad += compare_beams(db, {f: g['bor_' + f] for f in BEAM_FIELDS}, ('Es', 'Ep', 'Jss', 'Jpp', 'Jsp'), 1e-10, 'bor_')
    assert not bad, bad
    assert np.array_equal(dw['rDiffr'], g['rDiffr'])  # sqrt of numpy-ordered sums: the phase argument is the reference's
    for k in ('seeded', 'seededI', 'sourceWeight'):
        assert abs(getattr(bor, k) - float(g[k])) <= 1e-12 * abs(float(g[k])), k
    assert abs(bor.accepted - float(g['accepted'])) <= 1e-9 * float(g['accepted'])


@pytest.mark.gpu
def test_cuda_shine_wave_own_draws(cuda_device):
    """Without supplied deviates the source draws its own (torch Philox): a filament beam gives one energy and one
    source point for all samples, and the intensity on the screen is the undulator's central cone."""
    from xrt_b200.planes import Screen
    g = golden('shine_wave_filament')
    src, bl, E1 = _product_source(g)
    scr = Screen(bl, 'scr', center=list(g['screen_center']))
    wave = scr.prepare_wave(src, g['xm'], g['zm'])
    bor = src.shine(wave=wave)
    d = wave.to_numpy()
    assert np.unique(d['E']).size == 1 and float(g['E_min']) <= d['E'][0] <= float(g['E_max'])
    b = bor.to_numpy()
    assert np.unique(b['x']).size == 1 and np.all(d['state'] == 1)
    inten = (np.abs(d['Es'])**2 + np.abs(d['Ep'])**2).reshape(len(g['zm']), len(g['xm']))
    iz, ix = np.unravel_index(np.argmax(inten), inten.shape)
    assert abs(g['xm'][ix]) < 1.5 and abs(g['zm'][iz]) < 1.2 and np.isfinite(inten).all()


def test_oracle_shine_wave_accubeam_bit_exact():
    """shine(wave=wave2, accuBeam=bor): the same electron (energy, offsets, energy deviation) illuminates a second
    wave; seeded / seededI accumulate (sybase.py:1553-1561)."""
    g, src, wave, bor = _oracle_run('shine_wave_filament')
    for k in ('filamentDX', 'filamentDZ', 'filamentDtheta', 'filamentDpsi', 'filamentDgamma'):
        setattr(bor, k, getattr(wave, k))
    wave2 = osh.Wave(g['w2_xDiffr'], g['w2_yDiffr'], g['w2_zDiffr'], float(g['w2_area']))
    bor2 = osh.shine_wave(src, wave2, filamentBeam=True, accuBeam=bor)
    for f in ('a', 'b', 'c', 'E', 'Jss', 'Jpp', 'Jsp', 'Es', 'Ep', 'path', 'state'):
        assert np.array_equal(getattr(wave2, f), g['w2_' + f]), f
    assert np.array_equal(wave2.rDiffr, g['w2_rDiffr'])
    assert bor2.seeded == g['w2_seeded'] and bor2.seededI == g['w2_seededI']
    assert np.array_equal(bor2.x, g['w2_bor_x']) and np.array_equal(bor2.a, g['w2_bor_a'])


@pytest.mark.gpu
def test_cuda_shine_wave_accubeam(cuda_device):
    import torch
    from xrt_b200.planes import Screen
    from helpers import compare_beams
    rec = {}
    g, osrc, owave, obor = _oracle_run('shine_wave_filament', record=rec)
    src, bl, E1 = _product_source(g)
    src.Imax = float(g['Imax0'])
    src.fluxConst = src.Imax * src.xzE
    scr = Screen(bl, 'scr', center=list(g['screen_center']))
    wave = scr.prepare_wave(src, g['xm'], g['zm'])
    samples = {k: float(v) for k, v in rec.items() if k in ('uE', 'gX', 'gZ', 'gT', 'gP', 'gG')}
    bor = src.shine(wave=wave, samples=samples)
    wave2 = scr.prepare_wave(src, g['xm'][::2] + 0.3, g['zm'][::2] - 0.2)
    bor2 = src.shine(wave=wave2, accuBeam=bor)
    d = wave2.to_numpy()
    ref = {f: g['w2_' + f] for f in BEAM_FIELDS}
    geo = [f for f in BEAM_FIELDS if f not in ('Es', 'Ep', 'Jss', 'Jpp', 'Jsp', 'x', 'y', 'z')]
    bad = compare_beams(d, ref, geo, 1e-12, 'w2_')
    bad += compare_beams(d, ref, ('Es', 'Ep', 'Jss', 'Jpp', 'Jsp'), 1e-10, 'w2_')
    assert not bad, bad
    assert bor2.seeded == int(g['w2_seeded']) and abs(bor2.seededI - float(g['w2_seededI'])) <= 1e-12 * float(g['w2_seededI'])
    assert np.abs(bor2.to_numpy()['x'] - g['w2_bor_x']).max() <= 1e-12
