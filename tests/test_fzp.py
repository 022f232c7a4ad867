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


@pytest.mark.gpu
@pytest.mark.parametrize('name', CASES)
def test_cuda_fzp_vs_reference_golden(name, cuda_device):
    from xrt_b200 import oes, materials
    g = golden('reflect_' + name)
    kw = product_oe_kwargs(g)
    kw.pop('limPhysX'), kw.pop('limPhysY')
    oe = oes.NormalFZP(product_beamline(g), name, material=materials.Material('Au', rho=19.3, kind='FZP'),
                       f=float(g['f']), E=float(g['E0']), N=int(g['N']), isCentralZoneBlack=bool(g['black']),
                       order=int(g['order']), **kw)
    assert np.array_equal(oe.rn, g['rn']) and oe.limPhysX == list(g['surfPhysX'])
    oe.lostNum = int(g['lostNum'])
    gb, lb = oe.reflect(device_beam(oracle_beam(g, 'in_')))
    assert np.array_equal(gb.to_numpy()['state'], g['gb_state'])  # zone membership is exact
    bad = compare_beams(gb.to_numpy(), {f: g['gb_' + f] for f in OUT_FIELDS if 'gb_' + f in g.files}, OUT_FIELDS,
                        1e-12, 'gb_')
    bad += compare_beams(lb.to_numpy(), {f: g['lb_' + f] for f in OUT_FIELDS if 'lb_' + f in g.files}, OUT_FIELDS,
                         1e-12, 'lb_')
    assert not bad, bad
