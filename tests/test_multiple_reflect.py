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
