"""Edge cases of the elements added on top of the section-8 core (spliced surfaces, multiple_reflect, figure errors,
zone plate, round aperture): empty beams, beams without a single live ray, single rays -- the reference's behaviour is
'everything passes through untouched' (reflect.py:110-112, :198-201)."""
import numpy as np
import pytest

from conftest import golden
from helpers import oracle_beam, product_beamline, product_oe_kwargs, device_beam, BEAM_FIELDS


def _bent(g, **extra):
    from xrt_b200 import oes, materials
    kw = product_oe_kwargs(g)
    kw.update(extra)
    return oes.BentFlatMirror(product_beamline(g), 'vcm', R=float(g['R']),
                              material=materials.Material('Au', rho=float(g['rho'])), **kw)


@pytest.mark.gpu
def test_cuda_spliced_and_multi_all_dead_and_empty(cuda_device):
    g = golden('multi_reflect_bentflat')
    ob = oracle_beam(g, 'in_')
    ob.state[:] = -3
    beam = device_beam(ob)
    oe = _bent(g)
    gb, lb = oe.reflect(beam)
    d = gb.to_numpy()
    for f in BEAM_FIELDS:
        assert np.array_equal(d[f], getattr(ob, f)), f  # dead rays pass through untouched
    gbm, lbn = oe.multiple_reflect(beam, maxReflections=5)
    assert np.array_equal(gbm.to_numpy()['x'], ob.x) and len(lbn) == len(beam)
    # an empty beam
    from xrt_b200.beams import Beam
    empty = Beam(0, withAmplitudes=True, device=beam.device)
    gb0, lb0 = oe.reflect(empty)
    assert len(gb0) == 0 and len(lb0) == 0


@pytest.mark.gpu
def test_cuda_figure_error_single_ray_and_fzp_all_blocked(cuda_device):
    from xrt_b200 import oes, materials, figure_error as pfe
    g = golden('reflect_figerr_flat_waviness')
    ob = oracle_beam(g, 'in_')
    one = {f: getattr(ob, f)[5:6].copy() for f in BEAM_FIELDS}
    fe = pfe.FigureErrorMap(g['fe_x1d'], g['fe_y1d'], g['fe_z'])
    oe = oes.OE(product_beamline(g), 'flat', material=materials.Material('Au', rho=19.3), figureError=fe,
                **product_oe_kwargs(g))
    oe.lostNum = int(g['lostNum'])
    gb, _ = oe.reflect(device_beam(one))
    d = gb.to_numpy()
    assert d['state'][0] == g['gb_state'][5] and abs(d['x'][0] - g['gb_x'][5]) <= 1e-12 * 2e4
    # a zone plate much smaller than the beam: (almost) everything is absorbed, nothing is NaN
    g2 = golden('reflect_fzp_o1')
    kw = product_oe_kwargs(g2)
    kw.pop('limPhysX'), kw.pop('limPhysY')
    fzp = oes.NormalFZP(product_beamline(g2), 'tiny', material=materials.Material('Au', rho=19.3, kind='FZP'),
                        f=100., E=9000., N=3, **kw)
    out = fzp.reflect(device_beam(oracle_beam(g2, 'in_')))[0].to_numpy()
    assert (out['state'] == 1).sum() < 40 and np.isfinite(out['a']).all() and np.isfinite(out['Jss']).all()


@pytest.mark.gpu
def test_cuda_round_aperture_radius_zero_and_huge(cuda_device):
    from xrt_b200.planes import RoundAperture
    g = golden('propagate_round')
    for r, expect_good in ((0., False), (1e6, True)):
        ap = RoundAperture(None, 'pipe', center=list(g['center']), r=r)
        ap.x, ap.y, ap.z = list(g['ex']), list(g['ey']), list(g['ez'])
        ap.lostNum = int(g['lostNum'])
        ob = oracle_beam(g, 'in_')
        lo = ap.propagate(device_beam(ob)).to_numpy()
        alive = ob.state > 0
        assert np.all((lo['state'][alive] == 1) == expect_good)
        assert np.array_equal(lo['state'][~alive], ob.state[~alive])
