"""a13/a14: DCM.double_reflect with Si(111), symmetric and 5 deg asymmetric cut -- oracle bit-exact on the
reference's outputs (CPU), fused CUDA kernels vs golden (GPU)."""
import numpy as np
import pytest

from conftest import golden
from helpers import oracle_spec, oracle_beam, product_beamline, product_oe_kwargs, device_beam, compare_beams, \
    BEAM_FIELDS
from oracle import reflect as orf, materials as om

CASES = ['dcm_si111', 'dcm_si111_asym']
OUT_FIELDS = BEAM_FIELDS + ('theta',)
DCM_KEYS = ('bragg', 'cryst1roll', 'cryst2roll', 'cryst2pitch', 'cryst2finePitch', 'cryst2perpTransl',
            'cryst2longTransl')


@pytest.mark.parametrize('name', CASES)
def test_oracle_matches_reference_bit_exact(name):
    g = golden(name)
    si = om.CrystalDiamond((1, 1, 1), float(g['d']), 'Si')
    sp = oracle_spec(g, 'flat', si, alpha=float(g['alpha']) or None, material2=si,
                     surfPhysX2=list(g['surfPhysX2']), surfPhysY2=list(g['surfPhysY2']), surfOptX2=None,
                     surfOptY2=None, **{k: float(g[k]) for k in DCM_KEYS})
    gb, lo1, lo2 = orf.double_reflect(sp, oracle_beam(g, 'in_'))
    for pre, b in (('gb_', gb), ('lo1_', lo1), ('lo2_', lo2)):
        for f in OUT_FIELDS:
            if pre + f in g.files:
                assert np.array_equal(getattr(b, f), g[pre + f]), pre + f


@pytest.mark.gpu
@pytest.mark.parametrize('name', CASES)
def test_cuda_double_reflect_vs_golden(name, cuda_device):
    from xrt_b200 import oes, materials
    g = golden(name)
    si = materials.CrystalSi(hkl=(1, 1, 1), tK=float(g['tK']))
    assert si.d == float(g['d'])
    kw = product_oe_kwargs(g)
    kw.update({k: float(g[k]) for k in DCM_KEYS})
    dcm = oes.DCM(product_beamline(g), 'dcm', material=si, material2=si, alpha=float(g['alpha']) or None,
                  limPhysX2=list(g['surfPhysX2']), limPhysY2=list(g['surfPhysY2']), **kw)
    dcm.lostNum = int(g['lostNum'])
    gb, lo1, lo2 = dcm.double_reflect(device_beam(oracle_beam(g, 'in_')))
    bad = []
    for pre, b in (('gb_', gb), ('lo1_', lo1), ('lo2_', lo2)):
        ref = {f: g[pre + f] for f in OUT_FIELDS if pre + f in g.files}
        # the Bragg peak amplifies ulp-level differences of the transcendental functions like the Fresnel
        # case does: 1e-11 on the amplitude groups, 1e-12 on geometry
        bad += compare_beams(b.to_numpy(), ref, ('x', 'y', 'z', 'a', 'b', 'c', 'path', 'E', 'state', 'theta'),
                             1e-12, pre)
        bad += compare_beams(b.to_numpy(), ref, ('Jss', 'Jpp', 'Jsp', 'Es', 'Ep'), 1e-11, pre)
    assert not bad, bad
