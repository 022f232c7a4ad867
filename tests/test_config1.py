"""BASELINE.json configs[0], the reference's own CPU-runnable case, exactly as SURVEY.md 8d states it:
np.random.seed(1234); GeometricSource(1e5 rays) -> flat Au mirror (4 mrad) -> Screen.  The oracle chain reproduces the
reference's three beams bit for bit (every 50th ray and the whole-beam sums are in the fixture); the CUDA chain, fed with
the deviates the oracle drew, matches the oracle on all 1e5 rays."""
import numpy as np
import pytest

from conftest import golden
from oracle import geometric as og, reflect as orf, planes as op, materials as om

BEAM_FIELDS = ('x', 'y', 'z', 'a', 'b', 'c', 'path', 'E', 'Jss', 'Jpp', 'Jsp', 'state', 'Es', 'Ep')
SRC_KW = dict(nrays=100000, dx=0.1, dz=0.05, dxprime=1e-4, dzprime=5e-5, distE='flat', energies=(8990, 9010),
              polarization='h')
PITCH = 4e-3
OE_KW = dict(center=[0, 1e4, 0], pitch=PITCH)
LIM = dict(X=[-20, 20], Y=[-300, 300])
SCR_CENTER = [0, 2e4, 1e4*np.tan(2*PITCH)]


def _oracle_chain(record=None):
    g = golden('config1_chain')
    np.random.seed(int(g['seed']))
    src = og.shine(og.GeomSpec(**SRC_KW), withAmplitudes=True, record=record)
    oe = orf.OESpec(surface='flat', surfPhysX=LIM['X'], surfPhysY=LIM['Y'], material=om.Material('Au', rho=19.3),
                    lostNum=int(g['lostNum']), **OE_KW)
    info = {}
    gb, lb = orf.reflect(oe, src, info=info)
    blo = op.expose(gb, SCR_CENTER, [[1, 0, 0], [0, 1, 0], [0, 0, 1]], int(g['scr_lostNum']))
    return g, src, gb, lb, blo, info


def test_oracle_chain_matches_reference_bit_exact():
    g, src, gb, lb, blo, info = _oracle_chain()
    stride = int(g['stride'])
    for pre, b in (('src_', src), ('gb_', gb), ('lb_', lb), ('scr_', blo)):
        for f in BEAM_FIELDS:
            assert np.array_equal(getattr(b, f)[::stride], g[pre + f]), pre + f
            assert getattr(b, f).sum() == g[pre + f + '_sum'], pre + f
    assert np.array_equal(info['footprint'], g['footprint'])
    assert (blo.state == 1).sum() > 90000


@pytest.mark.gpu
def test_cuda_chain_vs_oracle_all_rays(cuda_device):
    import torch
    from helpers import compare_beams
    from xrt_b200.beamline import BeamLine
    from xrt_b200 import sources, oes, materials, planes
    rec = {}
    g, src_o, gb_o, lb_o, blo_o, info = _oracle_chain(rec)
    bl = BeamLine()
    src = sources.GeometricSource(bl, 'src', **SRC_KW)
    oe = oes.OE(bl, 'M1', material=materials.Material('Au', rho=19.3), limPhysX=LIM['X'], limPhysY=LIM['Y'], **OE_KW)
    scr = planes.Screen(bl, 'scr', center=SCR_CENTER)
    assert oe.lostNum == int(g['lostNum']) and scr.lostNum == int(g['scr_lostNum'])
    samples = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in rec.items() if v is not None}
    beam = src.shine(withAmplitudes=True, samples=samples)
    gb, lb = oe.reflect(beam, withFootprint=True)
    blo = scr.expose(gb)
    bad = []
    for name, dev, ref in (('src_', beam, src_o), ('gb_', gb, gb_o), ('lb_', lb, lb_o), ('scr_', blo, blo_o)):
        bad += compare_beams(dev.to_numpy(), ref, BEAM_FIELDS, 1e-12, name)
    assert not bad, bad
    np.testing.assert_allclose(oe.footprint[0], info['footprint'], rtol=1e-12, atol=1e-12)
