"""f4: figure errors -- OE(figureError=...) (oes/base.py:681-686, :744-770, find_dz :825-829, reflect.py:767-776;
figure_error.py:207-265): a height map (bicubic spline) added to the surface in the intersection search and its slopes
applied to the surface normal as d_pitch / d_roll rotations.  Oracle bit-exact on the reference's outputs (CPU); the
fused kernel with the spline evaluated on the device (GPU)."""
import numpy as np
import pytest

from conftest import golden
from helpers import (oracle_spec, oracle_beam, product_beamline, product_oe_kwargs, device_beam, compare_beams,
                     BEAM_FIELDS)
from oracle import materials as om, reflect as orf, surfaces as osf

CASES = [('toroid_bump', 'toroid'), ('flat_waviness', 'flat')]
OUT_FIELDS = BEAM_FIELDS + ('theta',)


@pytest.mark.parametrize('name,surface', CASES)
def test_oracle_figure_error_bit_exact(name, surface):
    g = golden('reflect_figerr_' + name)
    fe = osf.FigureErrorSpline(g['fe_x1d'], g['fe_y1d'], g['fe_z'], float(g['fe_xShift']), float(g['fe_yShift']))
    ty, tx = fe.local_z_spline.get_knots()
    assert np.array_equal(ty, g['fe_ty']) and np.array_equal(tx, g['fe_tx'])
    assert np.array_equal(fe.local_z_spline.get_coeffs(), g['fe_c'])
    kw = dict(surfParams=dict(R=float(g['R']), r=float(g['r']))) if surface == 'toroid' else {}
    sp = oracle_spec(g, surface, om.Material('Au', rho=float(g['rho'])), figureError=fe, **kw)
    info = {}
    gb, lb = orf.reflect(sp, oracle_beam(g, 'in_'), info=info)
    for pre, b in (('gb_', gb), ('lb_', lb)):
        for f in OUT_FIELDS:
            if pre + f in g.files:
                assert np.array_equal(getattr(b, f), g[pre + f]), pre + f
    assert np.array_equal(info['footprint'], g['footprint'])
