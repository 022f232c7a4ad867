"""f4: OE.multiple_reflect (oes/reflect.py:165-264) -- repeated reflections on one surface (whispering gallery): every
pass after the first leaves the surface through the root of the direction along the normal (derivOrder = 1,
base.py:1279-1289, :842-845) and then searches the next intersection.  Oracle bit-exact on the reference's outputs
(CPU); the CUDA passes against the same goldens (GPU)."""
import numpy as np
import pytest

from conftest import golden
from helpers import (oracle_spec, oracle_beam, product_beamline, product_oe_kwargs, device_beam, compare_beams,
                     BEAM_FIELDS)
from oracle import materials as om, reflect as orf, surfaces as osf

CASES = ['bentflat', 'toroid']
GB_FIELDS = BEAM_FIELDS + ('nRefl',)
LBN_FIELDS = BEAM_FIELDS + ('nRefl', 'theta')


def _oracle_spec(name, g):
    mat = om.Material('Au', rho=float(g['rho']))
    if name == 'toroid':
        return oracle_spec(g, 'toroid', mat, surfParams=dict(R=float(g['R']), r=float(g['r'])))
    return oracle_spec(g, 'custom', mat, surfParams={'obj': osf.BentFlatMirror(float(g['R']), float(g['surfPhysY'][0]))})


@pytest.mark.parametrize('name', CASES)
def test_oracle_multiple_reflect_bit_exact(name):
    g = golden('multi_reflect_' + name)
    infos = []
    gb, lbN = orf.multiple_reflect(_oracle_spec(name, g), oracle_beam(g, 'in_'), maxReflections=8, infos=infos)
    for pre, b, fields in (('gb_', gb, GB_FIELDS), ('lbN_', lbN, LBN_FIELDS)):
        for f in fields:
            assert np.array_equal(getattr(b, f), g[pre + f], equal_nan=True), pre + f
    assert len(lbN.x) == 8 * len(gb.x) == len(lbN.Es)
    assert np.array_equal(np.array([i['footprint'] for i in infos]), g['footprints'])
    assert gb.nRefl.max() == 8 and (gb.nRefl == 3).sum() > 20


ELEV = ('elevationD', 'elevationX', 'elevationY', 'elevationZ')


@pytest.mark.parametrize('name', CASES)
def test_oracle_elevation_map_bit_exact(name):
    """needElevationMap=True (reflect.py:212-216, :651-659; base.py:1285-1287): the highest point of every ray above the
    surface between two impacts (root of direction . normal) and its distance, per pass."""
    g = golden('multi_reflect_' + name)
    e = golden('multi_reflect_' + name + '_elevation')
    gb, lbN = orf.multiple_reflect(_oracle_spec(name, g), oracle_beam(g, 'in_'), maxReflections=8,
                                   needElevationMap=True)
    assert np.array_equal(gb.x, g['gb_x']) and np.array_equal(lbN.x, g['lbN_x'])  # the map does not change the rays
    for pre, b in (('gb_', gb), ('lbN_', lbN)):
        for f in ELEV:
            assert np.array_equal(getattr(b, f), e[pre + f]), pre + f
    n = len(gb.x)
    assert (lbN.elevationD[:n] == -1).all() and (lbN.elevationD[n:] > 0).sum() > 1000


def _product_oe(name, g):
    from xrt_b200 import oes, materials
    kw = product_oe_kwargs(g)
    kw['material'] = materials.Material('Au', rho=float(g['rho']))
    if name == 'toroid':
        oe = oes.ToroidMirror(product_beamline(g), name, R=float(g['R']), r=float(g['r']), **kw)
    else:
        oe = oes.BentFlatMirror(product_beamline(g), name, R=float(g['R']), **kw)  # a spliced surface (plug-in)
    oe.lostNum = int(g['lostNum'])
    return oe


@pytest.mark.gpu
@pytest.mark.parametrize('name', CASES)
def test_cuda_multiple_reflect_vs_reference_golden(name, cuda_device):
    """Every pass, every ray: states and reflection counts identical, geometry / coherency / amplitudes to 1e-12 of the
    group scales -- after 8 consecutive intersections, i.e. the solver iterates stay on the reference's (|dt| <= 2e-12
    mm per pass would already show up as 1e-4 rad in exp(i k t))."""
    g = golden('multi_reflect_' + name)
    oe = _product_oe(name, g)
    gb, lbN = oe.multiple_reflect(device_beam(oracle_beam(g, 'in_')), maxReflections=8)
    assert len(lbN) == len(g['lbN_x']) and len(oe.footprint) == len(g['footprints'])
    dgb, dlb = gb.to_numpy(), lbN.to_numpy()
    assert np.array_equal(dgb['nRefl'], g['gb_nRefl']) and np.array_equal(dlb['nRefl'], g['lbN_nRefl'])
    bad = compare_beams(dgb, {f: g['gb_' + f] for f in BEAM_FIELDS}, BEAM_FIELDS, 1e-12, 'gb_')
    bad += compare_beams(dlb, {f: g['lbN_' + f] for f in BEAM_FIELDS + ('theta',)}, BEAM_FIELDS + ('theta',), 1e-12,
                         'lbN_')
    assert not bad, bad
    np.testing.assert_allclose(np.array(oe.footprint), g['footprints'], rtol=1e-12, atol=1e-12)
    assert np.abs(dlb['path'] - g['lbN_path']).max() <= 2e-12 * 8


@pytest.mark.gpu
@pytest.mark.parametrize('name', CASES)
def test_cuda_elevation_map_vs_reference_golden(name, cuda_device):
    """multiple_reflect(needElevationMap=True): xrt_reflect_multi_elevation_f64 after the derivative stage of every pass
    (the built-in toroid and the spliced BentFlatMirror plug-in).  The map must not change the rays; the highest points
    follow the derivative roots (same iterates as the reference: 1e-9 mm), their heights are second order in them."""
    g = golden('multi_reflect_' + name)
    e = golden('multi_reflect_' + name + '_elevation')
    oe = _product_oe(name, g)
    beam = device_beam(oracle_beam(g, 'in_'))
    gb0, lbN0 = oe.multiple_reflect(beam, maxReflections=8)
    gb, lbN = oe.multiple_reflect(beam, maxReflections=8, needElevationMap=True)
    d0, d1 = lbN0.to_numpy(), lbN.to_numpy()
    for f in BEAM_FIELDS:
        assert np.array_equal(d0[f], d1[f], equal_nan=True), f
    for pre, b in (('gb_', gb), ('lbN_', lbN)):
        for f in ELEV:
            got = getattr(b, f).cpu().numpy()
            ref = e[pre + f]
            assert got.shape == ref.shape
            unset = ref == (-1. if f == 'elevationD' else -1000.)
            assert np.array_equal(got[unset], ref[unset]), pre + f  # rays without a map keep the initial values
            tol = 1e-12 if f == 'elevationD' else 1e-9
            assert np.abs(got - ref).max() <= tol, (pre + f, float(np.abs(got - ref).max()))
    assert (e['lbN_elevationD'] > 0).sum() > 1000


@pytest.mark.gpu
def test_cuda_multiple_reflect_single_pass_equals_reflect(cuda_device):
    """maxReflections = 1 is one reflect: the global beam of the good rays is OE.reflect's."""
    g = golden('multi_reflect_toroid')
    oe = _product_oe('toroid', g)
    beam = device_beam(oracle_beam(g, 'in_'))
    gb1, _ = oe.multiple_reflect(beam, maxReflections=1)
    gb2, _ = oe.reflect(beam)
    d1, d2 = gb1.to_numpy(), gb2.to_numpy()
    good = d2['state'] == 1
    assert good.sum() > 100 and np.array_equal(d1['state'] == 1, good)
    for f in BEAM_FIELDS:
        if f != 'state':
            assert np.array_equal(d1[f][good], d2[f][good]), f
