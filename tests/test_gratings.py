"""f4: gratings -- OE.reflect with material kind 'grating' (toWhere = 3, oes/reflect.py:840-861,
_grating_deflection :451-469, OE.local_g oes/base.py:688-717): plane VLS grating with Au coating, toroidal grating
without coating (EmptyMaterial, kind assigned from gratingDensity), constant line density with tabulated
efficiencies.  Oracle pinned bit-exact on the reference's outputs (CPU), CUDA kernel vs golden (GPU)."""
import numpy as np
import pytest

from conftest import golden
from helpers import (oracle_spec, oracle_beam, product_beamline, product_oe_kwargs, device_beam, compare_beams,
                     BEAM_FIELDS)
from oracle import reflect as orf, materials as om

CASES = [('reflect_grating_vls', 'flat'), ('reflect_grating_toroid', 'toroid'),
         ('reflect_grating_efficiency', 'flat')]
OUT_FIELDS = BEAM_FIELDS + ('theta',)


def _grating_kwargs(g):
    kw = dict(order=int(g['order']))
    if 'gratingDensity' in g.files:
        kw['gratingDensity'] = [str(g['gratingAxis'])] + [float(v) for v in g['gratingDensity']]
    return kw


def _efficiency(g):
    if 'eff_orders' in g.files:
        return [(int(o), float(v)) for o, v in zip(g['eff_orders'], g['eff_values'])]
    return None


def _oracle_run(name, surface):
    g = golden(name)
    if 'rho' in g.files:
        mat = om.Material('Au', rho=float(g['rho']), kind='grating', efficiency=_efficiency(g))
    else:
        mat = om.EmptyMaterial()
    kw = _grating_kwargs(g)
    if surface == 'toroid':
        kw['surfParams'] = dict(R=float(g['R']), r=float(g['r']))
    info = {}
    gb, lb = orf.reflect(oracle_spec(g, surface, mat, **kw), oracle_beam(g, 'in_'), info=info)
    return g, gb, lb, info


@pytest.mark.parametrize('name,surface', CASES)
def test_oracle_matches_reference_bit_exact(name, surface):
    g, gb, lb, info = _oracle_run(name, surface)
    for pre, b in (('gb_', gb), ('lb_', lb)):
        for f in OUT_FIELDS:
            if pre + f in g.files:
                assert np.array_equal(getattr(b, f), g[pre + f]), pre + f
    assert np.array_equal(info['footprint'], g['footprint'])
    st = g['gb_state']
    assert (st == 1).sum() > 50 and (st < 0).sum() > 10


def _product_oe(g, surface):
    from xrt_b200 import oes, materials
    bl = product_beamline(g)
    kw = product_oe_kwargs(g)
    kw.update(_grating_kwargs(g))
    if 'rho' in g.files:
        # 'auto' becomes 'grating' through the OE's gratingDensity (base.py:1089-1093); without one the kind is given
        kind = 'auto' if 'gratingDensity' in g.files else 'grating'
        kw['material'] = materials.Material('Au', rho=float(g['rho']), kind=kind, efficiency=_efficiency(g))
    if surface == 'toroid':
        oe = oes.ToroidMirror(bl, 'oe', R=float(g['R']), r=float(g['r']), **kw)
    else:
        oe = oes.OE(bl, 'oe', **kw)
    oe.lostNum = int(g['lostNum'])
    return oe


# The grating branch adds pow() (VLS terms beyond the quadratic one) and one more sqrt / division chain to the
# mirror kernel; everything is evaluated in the reference's operation order: same 1e-12 bar as test_reflect.py.
RTOL = 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize('name,surface', CASES)
def test_cuda_grating_vs_reference_golden(name, surface, cuda_device):
    g = golden(name)
    oe = _product_oe(g, surface)
    gb, lb = oe.reflect(device_beam(oracle_beam(g, 'in_')), withFootprint=True)
    bad = []
    for pre, b in (('gb_', gb), ('lb_', lb)):
        ref = {f: g[pre + f] for f in OUT_FIELDS if pre + f in g.files}
        bad += compare_beams(b.to_numpy(), ref, OUT_FIELDS, RTOL, pre)
    assert not bad, bad
    np.testing.assert_allclose(oe.footprint[0], g['footprint'], rtol=1e-12, atol=1e-12)
    good = g['lb_state'] == 1
    assert np.abs(lb.to_numpy()['path'][good] - g['lb_path'][good]).max() <= 2e-12


@pytest.mark.gpu
def test_cuda_grating_order_zero_is_specular(cuda_device):
    """Order 0 leaves the mirror direction: a - 2 (a.n) n, whatever the line density."""
    from xrt_b200 import oes
    g = golden('reflect_grating_vls')
    kw = product_oe_kwargs(g)
    beam = device_beam(oracle_beam(g, 'in_'))
    grating = oes.OE(product_beamline(g), 'g', gratingDensity=['y', 1200., 1.], order=0, **kw)
    mirror = oes.OE(product_beamline(g), 'm', **kw)
    gb1, lb1 = grating.reflect(beam)
    gb2, lb2 = mirror.reflect(beam)
    a, b = lb1.to_numpy(), lb2.to_numpy()
    assert np.array_equal(a['state'], b['state'])
    good = a['state'] == 1
    for f in ('a', 'b', 'c'):
        assert np.abs(a[f][good] - b[f][good]).max() < 1e-15


def test_order_sequence_is_refused():
    from xrt_b200 import oes
    oe = oes.OE(None, 'g', gratingDensity=['y', 600., 1.], order=(1, 2))
    with pytest.raises(NotImplementedError):
        oe._make_params(0., 0., 0.)
