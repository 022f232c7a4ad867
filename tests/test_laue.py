"""f2 / f4: Laue-geometry elements -- LauePlate (flat; oes/laue.py:11-24) and BentLaueCylinder (oes/laue.py:26-175)
with Si(111) 'Laue reflected', with and without asymmetry angle: the planes' normal and the surface normal are
returned together.  Oracle bit-exact on the reference's outputs (CPU); CUDA against the same goldens (GPU): LauePlate
through the built-in flat kernel, BentLaueCylinder through its spliced surface."""
import numpy as np
import pytest

from conftest import golden
from helpers import (oracle_spec, oracle_beam, product_beamline, product_oe_kwargs, device_beam, compare_beams,
                     BEAM_FIELDS)
from oracle import materials as om, reflect as orf, surfaces as osf

CASES = ['laueplate', 'laueplate_alpha', 'bentlaue_parab', 'bentlaue_circ_alpha']
OUT_FIELDS = BEAM_FIELDS + ('theta',)


@pytest.mark.parametrize('name', CASES)
def test_oracle_laue_elements_bit_exact(name):
    g = golden('reflect_laue_' + name)
    alpha = float(g['alpha']) or None
    obj = osf.LauePlate(alpha) if name.startswith('laueplate') else \
        osf.BentLaueCylinder(float(g['R']), 'circular' if 'circ' in name else 'parabolic', alpha)
    si = om.CrystalDiamond((1, 1, 1), float(g['d']), 'Si', t=float(g['t']), geom='Laue reflected')
    sp = oracle_spec(g, 'custom', si, surfParams={'obj': obj}, alpha=alpha)
    gb, lb = orf.reflect(sp, oracle_beam(g, 'in_'))
    for pre, b in (('gb_', gb), ('lb_', lb)):
        for f in OUT_FIELDS:
            if pre + f in g.files:
                assert np.array_equal(getattr(b, f), g[pre + f]), pre + f
    good = g['gb_state'] == 1
    assert 0.03 < (g['gb_Jss'] + g['gb_Jpp'])[good].mean() / (g['in_Jss'] + g['in_Jpp'])[good].mean() < 0.2


@pytest.mark.gpu
@pytest.mark.parametrize('name', CASES)
def test_cuda_laue_elements_vs_reference_golden(name, cuda_device):
    from xrt_b200 import oes
    from xrt_b200.materials import CrystalSi
    g = golden('reflect_laue_' + name)
    alpha = float(g['alpha']) or None
    si = CrystalSi(hkl=(1, 1, 1), geom='Laue reflected', t=float(g['t']), tK=float(g['tK']))
    kw = dict(material=si, alpha=alpha, **product_oe_kwargs(g))
    if name.startswith('laueplate'):
        oe = oes.LauePlate(product_beamline(g), name, **kw)
        assert oe._plugin() is None  # flat: the library's own kernel
    else:
        oe = oes.BentLaueCylinder(product_beamline(g), name, R=float(g['R']),
                                  crossSection='circular' if 'circ' in name else 'parabolic', **kw)
    oe.lostNum = int(g['lostNum'])
    gb, lb = oe.reflect(device_beam(oracle_beam(g, 'in_')))
    bad = compare_beams(gb.to_numpy(), {f: g['gb_' + f] for f in OUT_FIELDS if 'gb_' + f in g.files}, OUT_FIELDS,
                        1e-11, 'gb_')
    bad += compare_beams(lb.to_numpy(), {f: g['lb_' + f] for f in OUT_FIELDS if 'lb_' + f in g.files}, OUT_FIELDS,
                         1e-11, 'lb_')
    assert not bad, bad
    assert np.array_equal(gb.to_numpy()['state'], g['gb_state'])
