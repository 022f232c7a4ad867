"""a6-a12: OE.reflect (flat / toroid, Au mirror or bare) -- oracle pinned bit-exact on the reference's outputs
(CPU) and the fused CUDA kernel against golden + oracle (GPU)."""
import numpy as np
import pytest

from conftest import golden
from helpers import (oracle_spec, oracle_beam, product_beamline, product_oe_kwargs, device_beam, compare_beams,
                     BEAM_FIELDS)
from oracle import reflect as orf, materials as om

CASES = [('reflect_flat_Au', 'flat'), ('reflect_flat_round', 'flat'),  # shape='round' (base.py:1140-1156)
         ('reflect_toroid_Au', 'toroid'), ('reflect_toroid_bare', 'toroid'),
         ('reflect_toroid_brent', 'toroid'),  # drives _use_Brent_method (base.py:871, :961-1048)
         ('reflect_ellipse_vfm', 'ellipse_param'), ('reflect_ellipse_hfm', 'ellipse_param'),
         ('reflect_ellipsoid', 'ellipse_param')]
OUT_FIELDS = BEAM_FIELDS + ('theta', 's', 'phi', 'r')
ELLIPSE_KEYS = ('ellipseA', 'ellipseB', 'y0', 'z0', 'cosGamma', 'sinGamma')


def _oracle_run(name, surface):
    g = golden(name)
    mat = om.Material('Pt' if 'brent' in name else 'Au', rho=float(g['rho'])) if 'rho' in g.files else None
    kw = {}
    if surface == 'toroid':
        kw['surfParams'] = dict(R=float(g['R']), r=float(g['r']))
    if surface == 'ellipse_param':
        kw['isParametric'] = True
        kw['surfParams'] = dict(isCylindrical=bool(g['isCylindrical']), **{k: float(g[k]) for k in ELLIPSE_KEYS})
    if 'round' in name:
        kw['shape'] = 'round'
    sp = oracle_spec(g, surface, mat, **kw)
    info = {}
    gb, lb = orf.reflect(sp, oracle_beam(g, 'in_'), info=info)
    return g, gb, lb, info


@pytest.mark.parametrize('name,surface', CASES)
def test_oracle_matches_reference_bit_exact(name, surface):
    g, gb, lb, info = _oracle_run(name, surface)
    for pre, b in (('gb_', gb), ('lb_', lb)):
        for f in OUT_FIELDS:
            if pre + f in g.files:
                assert np.array_equal(getattr(b, f), g[pre + f]), pre + f
    assert np.array_equal(info['footprint'], g['footprint'])
    assert info['useBrent'] == ('brent' in name)
    # the goldens exercise every state class
    st = g['gb_state']
    assert (st == 1).sum() > 50 and (st < 0).sum() > 10


def _product_oe(g, surface, name=''):
    from xrt_b200 import oes, materials
    bl = product_beamline(g)
    kw = product_oe_kwargs(g)
    if 'round' in name:
        kw['shape'] = 'round'
    if 'rho' in g.files:
        kw['material'] = materials.Material('Pt' if float(g['rho']) > 20 else 'Au', rho=float(g['rho']))
    if surface == 'toroid':
        oe = oes.ToroidMirror(bl, 'oe', R=float(g['R']), r=float(g['r']), **kw)
    elif surface == 'ellipse_param':
        oe = oes.EllipticalMirrorParam(bl, 'oe', p=float(g['p']), q=float(g['q']),
                                       isCylindrical=bool(g['isCylindrical']), **kw)
        # the host mirror derives the ellipse from p, q, pitch like EllipticalMirrorParam._reset_pq
        for k in ELLIPSE_KEYS:
            assert abs(getattr(oe, k) - float(g[k])) <= 1e-13 * max(1., abs(float(g[k]))), k
            setattr(oe, k, float(g[k]))
    else:
        oe = oes.OE(bl, 'oe', **kw)
    oe.lostNum = int(g['lostNum'])
    return oe


# Tolerances: the kernel evaluates the reference's operation sequence without FMA contraction,
# so geometry is expected bit-identical up to libm-vs-CUDA differences of transcendental functions (<= 2 ulp in
# normals / angles): positions 1e-12 relative to the 1e4 mm scale, directions and amplitudes 1e-12.
RTOL = 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize('name,surface', CASES)
def test_cuda_reflect_vs_reference_golden(name, surface, cuda_device):
    from xrt_b200.physconsts import CHBAR
    g = golden(name)
    oe = _product_oe(g, surface, name)
    beam = device_beam(oracle_beam(g, 'in_'))
    gb, lb = oe.reflect(beam, withFootprint=True)
    bad = []
    good = g['lb_state'] == 1
    for pre, b in (('gb_', gb), ('lb_', lb)):
        ref = {f: g[pre + f] for f in OUT_FIELDS if pre + f in g.files}
        dev = b.to_numpy()
        if surface == 'ellipse_param' and 'Es' in dev:
            # The parametric surface goes through cos/atan2, where CUDA and glibc differ by an ulp; the secant
            # iteration then stops at a t that is 1-2 ulp (<= 4e-12 mm) away -- inside the reference's own
            # zEps/sin(pitch) ~ 2.5e-10 mm indeterminacy -- and exp(i k t) turns that into ~2e-4 rad.  Compare the
            # amplitudes with that propagation phase taken out, and bound the path difference itself.
            dpath = dev['path'] - ref['path']
            assert np.abs(dpath[good]).max() <= 4e-12
            comp = np.exp(-1e7j * ref['E'] / CHBAR * dpath)
            dev = dict(dev, Es=dev['Es'] * comp, Ep=dev['Ep'] * comp)
            # what is left is the rounding of the phase k*t itself (~1.4e12 rad, ulp 2.4e-4 rad) at two
            # different t: bound it by 2 ulp of the largest phase
            phase_ulp = np.spacing((1e7 * ref['E'] / CHBAR * ref['path']).max())
            bad += compare_beams(dev, ref, ('Es', 'Ep'), 2 * phase_ulp, pre)
            bad += compare_beams(dev, ref, [f for f in OUT_FIELDS if f not in ('Es', 'Ep')], RTOL, pre)
            continue
        bad += compare_beams(dev, ref, OUT_FIELDS, RTOL, pre)
    assert not bad, bad
    np.testing.assert_allclose(oe.footprint[0], g['footprint'], rtol=1e-12, atol=1e-12)
    if surface != 'ellipse_param':
        # intersection parameter: |dt| <= 2e-12 mm -- bit-identical iterates in practice
        assert np.abs(lb.to_numpy()['path'][good] - g['lb_path'][good]).max() <= 2e-12


@pytest.mark.gpu
@pytest.mark.parametrize('name,surface', [CASES[0], CASES[2]])
def test_cuda_reflect_need_local_false(name, surface, cuda_device):
    g = golden(name)
    oe = _product_oe(g, surface)
    gb, lb = oe.reflect(device_beam(oracle_beam(g, 'in_')), needLocal=False)
    assert lb is gb
    ref = {f: g['gb_' + f] for f in BEAM_FIELDS if 'gb_' + f in g.files}
    assert not compare_beams(gb.to_numpy(), ref, BEAM_FIELDS, RTOL)


@pytest.mark.gpu
def test_cuda_reflect_all_dead_and_empty(cuda_device):
    g = golden('reflect_flat_Au')
    oe = _product_oe(g, 'flat')
    b = oracle_beam(g, 'in_')
    b.state[:] = -3
    gb, lb = oe.reflect(device_beam(b))
    out = gb.to_numpy()
    for f in BEAM_FIELDS:
        assert np.array_equal(out[f], getattr(b, f)), f
    # zero-length beam
    from xrt_b200.beams import Beam
    e = Beam(0, withAmplitudes=True)
    gb, lb = oe.reflect(e)
    assert len(gb) == 0 and len(lb) == 0


@pytest.mark.gpu
def test_cuda_reflect_energy_outside_table_raises(cuda_device):
    g = golden('reflect_flat_Au')
    oe = _product_oe(g, 'flat')
    b = oracle_beam(g, 'in_')
    b.E[:] = 1e6  # above the Chantler table (4.3e5 eV): Element.get_f1f2 raises ValueError (element.py:254)
    with pytest.raises(ValueError):
        oe.reflect(device_beam(b))


@pytest.mark.gpu
@pytest.mark.parametrize('surface', ['flat', 'toroid'])
def test_cuda_reflect_vs_oracle_large_seeded(surface, cuda_device):
    """2e5 seeded rays (oracle finishes in seconds), ragged length, same comparison against the oracle."""
    from xrt_b200 import synthetic
    g = golden('reflect_toroid_Au' if surface == 'toroid' else 'reflect_flat_Au')
    n = 200003
    arrays = synthetic.geometric_beam(41, n, dx=0.3, dz=0.2, dxprime=1.5e-4, dzprime=8e-5,
                                      energies=(7000., 11000.), azimuth=np.arctan2(float(g['sinAzimuth']),
                                                                                   float(g['cosAzimuth'])))
    ob_beam = oracle_beam({'in_' + k: v for k, v in arrays.items()}, 'in_') if False else None
    from oracle import beam as obm
    b = obm.Beam.__new__(obm.Beam)
    for k, v in arrays.items():
        setattr(b, k, v.copy())
    mat = om.Material('Au', rho=float(g['rho']))
    kw = {'surfParams': dict(R=float(g['R']), r=float(g['r']))} if surface == 'toroid' else {}
    gbo, lbo = orf.reflect(oracle_spec(g, surface, mat, **kw), b)
    oe = _product_oe(g, surface)
    gb, lb = oe.reflect(device_beam(arrays))
    bad = compare_beams(gb.to_numpy(), gbo, BEAM_FIELDS, RTOL, 'gb_') + \
        compare_beams(lb.to_numpy(), lbo, OUT_FIELDS, RTOL, 'lb_')
    assert not bad, bad
