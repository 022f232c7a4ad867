"""GeometricSource.shine (sources/geoms.py:193-534), the input generator of BASELINE.json configs 1 and 3: oracle
pinned bit-exact on the reference's seeded outputs (CPU); the CUDA source kernel fed with the oracle's deviates
against the golden, and statistical checks of its own Philox sampling (GPU)."""
import numpy as np
import pytest

from conftest import golden
from oracle import geometric as og

BEAM_FIELDS = ('x', 'y', 'z', 'a', 'b', 'c', 'path', 'E', 'Jss', 'Jpp', 'Jsp', 'state', 'Es', 'Ep')

# the same keyword sets tools/gen_golden.py gave to the reference
CASES = {
    'geom_config1': dict(seed=1234, withAmplitudes=True, kw=dict(
        nrays=4000, dx=0.1, dz=0.05, dxprime=1e-4, dzprime=5e-5, distE='flat', energies=(8990, 9010),
        polarization='h')),
    'geom_uniform_annulus': dict(seed=77, withAmplitudes=True, kw=dict(
        nrays=3000, center=(1., -2., 0.5), distx='flat', dx=(-0.3, 0.5), distz='normal', dz=(0.02, 0.07),
        disty='flat', dy=2., distxprime='annulus', dxprime=(1e-4, 3e-4), distzprime='annulus', dzprime=(0.3, 2.5),
        distE='lines', energies=(8000., 9000., 12000.), energyWeights=(0.2, 0.5, 0.3), polarization='r',
        uniformRayDensity=True, pitch=1e-3, roll=0.3, yaw=-2e-3)),
    'geom_unpolarized_flux': dict(seed=78, withAmplitudes=True, kw=dict(
        nrays=3000, disty='normal', dy=5., distE='normal', energies=(9000., 3.), polarization=None,
        totalFlux=1e12)),
    'geom_custom_pol': dict(seed=79, withAmplitudes=False, kw=dict(
        nrays=3000, distz=None, dz=0, distE='lines', energies=(5000., 7000.), polarization=(0.7, 0.3, 0.1, -0.2),
        dxprime=2e-3, dzprime=1.5e-3)),
}


def _oracle(name, record=None):
    g = golden(name)
    c = CASES[name]
    spec = og.GeomSpec(azimuth=(float(g['sinAzimuth']), float(g['cosAzimuth'])), **c['kw'])
    np.random.seed(c['seed'])
    return g, spec, og.shine(spec, withAmplitudes=c['withAmplitudes'], record=record)


@pytest.mark.parametrize('name', list(CASES))
def test_oracle_matches_reference_bit_exact(name):
    g, spec, beam = _oracle(name)
    for f in BEAM_FIELDS:
        if 'out_' + f in g.files:
            assert np.array_equal(getattr(beam, f), g['out_' + f]), f
        else:
            assert not hasattr(beam, f), f
    for k in ('sourceWeight', 'seeded', 'seededI', 'accepted', 'acceptedE'):
        if k in g.files:
            assert getattr(beam, k) == g[k], k


# ------------------------------------------------------------------ GPU
def _product(name, g, **over):
    from xrt_b200 import sources
    from helpers import product_beamline
    kw = dict(CASES[name]['kw'])
    kw.update(over)
    return sources.GeometricSource(product_beamline(g), 'gs', **kw)


# Rounding order of the reference throughout; exp (uniformRayDensity weights) and cos / sin (annulus) differ from
# glibc by <= 2 ulp: 1e-12 of the group scales.
@pytest.mark.gpu
@pytest.mark.parametrize('name', list(CASES))
def test_cuda_shine_vs_reference_golden_with_shared_deviates(name, cuda_device):
    import torch
    from helpers import compare_beams
    rec = {}
    g, spec, beam = _oracle(name, rec)
    src = _product(name, g)
    samples = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in rec.items() if v is not None}
    out = src.shine(withAmplitudes=CASES[name]['withAmplitudes'], samples=samples)
    dev = out.to_numpy()
    ref = {f: g['out_' + f] for f in BEAM_FIELDS if 'out_' + f in g.files}
    bad = compare_beams(dev, ref, BEAM_FIELDS, 1e-12)
    assert not bad, bad
    assert ('Es' in dev) == ('out_Es' in g.files)
    if 'sourceWeight' in g.files:
        assert abs(out.sourceWeight - float(g['sourceWeight'])) <= 1e-12 * float(g['sourceWeight'])
        assert out.seeded == int(g['seeded'])


@pytest.mark.gpu
@pytest.mark.parametrize('name', list(CASES))
def test_cuda_shine_philox_statistics(name, cuda_device):
    """The kernel's own draws: moments of every field against a large numpy draw of the oracle."""
    g = golden(name)
    n = 400000
    src = _product(name, g, nrays=n)
    a = src.shine(withAmplitudes=CASES[name]['withAmplitudes']).to_numpy()
    kw = dict(CASES[name]['kw'], nrays=n)
    spec = og.GeomSpec(azimuth=(float(g['sinAzimuth']), float(g['cosAzimuth'])), **kw)
    np.random.seed(5)
    b = og.shine(spec, withAmplitudes=CASES[name]['withAmplitudes']).__dict__
    assert np.all(a['state'] == 1) and np.abs(a['a']**2 + a['b']**2 + a['c']**2 - 1).max() < 1e-14
    for f in ('x', 'y', 'z', 'a', 'c', 'E', 'Jss', 'Jpp'):
        sa, sb = a[f].std(), b[f].std()
        err = 5 * np.sqrt((sa**2 + sb**2) / n) + 1e-14 * (1 + abs(b[f].mean()))
        assert abs(a[f].mean() - b[f].mean()) < err, (f, a[f].mean(), b[f].mean())
        assert abs(sa - sb) <= 0.02 * sb + 1e-14 * (1 + abs(b[f].mean())), (f, sa, sb)
    if 'lines' == CASES[name]['kw'].get('distE'):
        va, ca = np.unique(a['E'], return_counts=True)
        vb, cb = np.unique(b['E'], return_counts=True)
        assert np.array_equal(va, vb) and np.abs(ca / n - cb / n).max() < 0.01
    # two calls never reuse deviates; two sources with the same seed reproduce each other
    a2 = src.shine(withAmplitudes=CASES[name]['withAmplitudes']).to_numpy()
    assert not np.array_equal(a2['x'], a['x']) or CASES[name]['kw'].get('distx') is None
    src2 = _product(name, g, nrays=n)
    a3 = src2.shine(withAmplitudes=CASES[name]['withAmplitudes']).to_numpy()
    assert np.array_equal(a3['x'], a['x']) and np.array_equal(a3['E'], a['E'])
