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
