"""f4: NormalFZP (oes/gratings.py:10-137) with a material of kind 'FZP' (toWhere = 4, reflect.py:745-746, :840-861):
zone mask (opaque zones absorb) + radial grating deflection, amplitude 1.  Oracle bit-exact on the reference's outputs
(CPU); the fused kernel against the same goldens (GPU)."""
import numpy as np
import pytest

from conftest import golden
from helpers import (oracle_spec, oracle_beam, product_beamline, product_oe_kwargs, device_beam, compare_beams,
                     BEAM_FIELDS)
from oracle import materials as om, reflect as orf, surfaces as osf

CASES = ['fzp_o1', 'fzp_o3_white']
OUT_FIELDS = BEAM_FIELDS + ('theta',)


@pytest.mark.parametrize('name', CASES)
def test_oracle_fzp_bit_exact(name):
    g = golden('reflect_' + name)
    fzp = osf.NormalFZP(float(g['f']), float(g['E0']), int(g['N']), bool(g['black']))
    assert np.array_equal(fzp.rn, g['rn'])
    mat = om.Material('Au', rho=19.3)
    mat.kind = 'FZP'
    sp = oracle_spec(g, 'flat', mat, fzp=fzp, order=int(g['order']))
    gb, lb = orf.reflect(sp, oracle_beam(g, 'in_'))
    for pre, b in (('gb_', gb), ('lb_', lb)):
        for f in OUT_FIELDS:
            if pre + f in g.files:
                assert np.array_equal(getattr(b, f), g[pre + f]), pre + f
    assert (g['gb_state'] == 1).sum() > 1000 and (g['gb_state'] == int(g['lostNum'])).sum() > 1000
