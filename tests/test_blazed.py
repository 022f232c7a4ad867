"""f4: BlazedGrating with a constant line density (oes/gratings.py:316-535): sawtooth surface, its own analytic
find_intersection (first facet hit), facet normals, material kind 'mirror'.  Oracle bit-exact on the reference's
outputs (CPU); the fused kernel with the analytic intersection in place of the solver (GPU)."""
import numpy as np
import pytest

from conftest import golden
from helpers import (oracle_spec, oracle_beam, product_beamline, product_oe_kwargs, device_beam, compare_beams,
                     BEAM_FIELDS)
from oracle import materials as om, reflect as orf, surfaces as osf

CASES = ['blazed', 'blazed_anti']
OUT_FIELDS = BEAM_FIELDS + ('theta',)


@pytest.mark.parametrize('name', CASES)
def test_oracle_blazed_grating_bit_exact(name):
    g = golden('reflect_' + name)
    obj = osf.BlazedGrating(float(g['blaze']), float(g['antiblaze']), float(g['rho0']))
    sp = oracle_spec(g, 'custom', om.Material('Au', rho=float(g['rho'])), surfParams={'obj': obj})
    gb, lb = orf.reflect(sp, oracle_beam(g, 'in_'))
    for pre, b in (('gb_', gb), ('lb_', lb)):
        for f in OUT_FIELDS:
            if pre + f in g.files:
                assert np.array_equal(getattr(b, f), g[pre + f]), pre + f
    assert (g['gb_state'] == 1).sum() > 2000


@pytest.mark.gpu
@pytest.mark.parametrize('name', CASES)
def test_cuda_blazed_grating_vs_reference_golden(name, cuda_device):
    from xrt_b200 import oes, materials
    g = golden('reflect_' + name)
    oe = oes.BlazedGrating(product_beamline(g), name, material=materials.Material('Au', rho=float(g['rho'])),
                           blaze=float(g['blaze']), antiblaze=float(g['antiblaze']), rho=float(g['rho0']),
                           **product_oe_kwargs(g))
    oe.lostNum = int(g['lostNum'])
    assert oe._plugin().own_intersection
    gb, lb = oe.reflect(device_beam(oracle_beam(g, 'in_')), withFootprint=True)
    assert np.array_equal(gb.to_numpy()['state'], g['gb_state'])
    # the facet a ray lands on is decided by floor / fmod of exactly the reference's operands: identical facets,
    # everything else 1e-12 (Es / Ep carry exp(i k t) with t from the same closed-form expressions)
    bad = compare_beams(gb.to_numpy(), {f: g['gb_' + f] for f in OUT_FIELDS if 'gb_' + f in g.files}, OUT_FIELDS,
                        1e-12, 'gb_')
    bad += compare_beams(lb.to_numpy(), {f: g['lb_' + f] for f in OUT_FIELDS if 'lb_' + f in g.files}, OUT_FIELDS,
                         1e-12, 'lb_')
    assert not bad, bad
    np.testing.assert_allclose(oe.footprint[0], g['footprint'], rtol=1e-12, atol=1e-12)
