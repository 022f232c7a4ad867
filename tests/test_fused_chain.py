"""Chain fusion (r02): OE.reflect / DCM.double_reflect continued by Screen.expose in ONE kernel
(xrt_reflect_expose_f64).  The contract is bit-for-bit equality with the two-call path, which is parity-tested against
the reference's goldens (test_reflect.py, test_dcm.py, test_planes.py): the fused kernel runs the same device code and
only skips the round trip of the global beam through HBM."""
import numpy as np
import pytest

from conftest import golden
from helpers import oracle_beam, device_beam
from test_reflect import _product_oe

CASES = [('reflect_flat_Au', 'flat'), ('reflect_flat_round', 'flat'), ('reflect_toroid_Au', 'toroid'),
         ('reflect_toroid_bare', 'toroid'), ('reflect_toroid_brent', 'toroid'), ('reflect_ellipse_vfm', 'ellipse_param'),
         ('reflect_ellipse_hfm', 'ellipse_param')]


def _same(a, b, what):
    import torch
    ta, tb = a.tensors(), b.tensors()
    assert set(ta) == set(tb), what
    for k, v in ta.items():
        w = tb[k]
        if v.is_complex():
            v, w = torch.view_as_real(v), torch.view_as_real(w)
        # NaN-safe bitwise comparison
        assert bool((v.contiguous().view(torch.int64 if v.element_size() == 8 else torch.int32) ==
                     w.contiguous().view(torch.int64 if w.element_size() == 8 else torch.int32)).all()), (what, k)


def _screen(oe, g, tilt=False):
    from xrt_b200.planes import Screen
    c = np.asarray(g['center'], dtype=float)
    kw = {}
    if tilt:
        kw = dict(x=[np.cos(0.1), np.sin(0.1), 0.], z=[0., 0., 1.], compressX=0.5)
    return Screen(oe.bl, 'scr', center=[c[0], c[1] + 1500., c[2] + 3.], **kw)


@pytest.mark.gpu
@pytest.mark.parametrize('name,surface', CASES)
@pytest.mark.parametrize('needLocal', [True, False])
def test_reflect_expose_fused_equals_two_calls(name, surface, needLocal, cuda_device):
    g = golden(name)
    oe = _product_oe(g, surface, name)
    scr = _screen(oe, g, tilt=(surface == 'toroid'))
    beam = device_beam(oracle_beam(g, 'in_'))
    gb, lb = oe.reflect(beam, needLocal=needLocal)
    ref = scr.expose(gb, withFootprint=True)
    foot_ref = scr.footprint[0].copy()
    out, lbf = oe.reflect(beam, needLocal=needLocal, exposeOn=scr)
    _same(out, ref, name)
    if needLocal:
        _same(lbf, lb, name + ' lb')
    # the screen's footprint slots are filled like xrt_expose_f64 fills them
    foot = scr._decode_minmax(scr._red(beam.device))
    assert np.array_equal(foot, foot_ref)
    st = ref.to_numpy()['state']
    assert (st == 1).sum() > 50 and (st != 1).sum() > 10


@pytest.mark.gpu
@pytest.mark.parametrize('n', [1, 255, 256, 257, 1000, 4099])
def test_reflect_expose_fused_ragged_sizes_and_views(n, cuda_device):
    """Tile edges: the last (ragged) tile of the bulk-copy staging, and beams that are views at odd offsets (no
    16-byte alignment: the kernel stages the field tile with plain loads)."""
    from xrt_b200.beams import Beam
    g = golden('reflect_toroid_Au')
    oe = _product_oe(g, 'toroid', 'reflect_toroid_Au')
    scr = _screen(oe, g)
    full = oracle_beam(g, 'in_')
    d = {k: v for k, v in full.__dict__.items() if isinstance(v, np.ndarray)}
    reps = -(-(n + 1) // len(d['x']))
    big = {k: np.concatenate([v] * reps) for k, v in d.items()}
    whole = Beam.from_numpy({k: v[1:n + 1].copy() for k, v in big.items()})     # aligned allocation
    dev_big = Beam.from_numpy(big)
    view = Beam.from_numpy({k: v[:1].copy() for k, v in big.items()})           # placeholder of the right type
    for k, t in dev_big.tensors().items():
        setattr(view, k, t[1:n + 1])                                             # a view starting at element 1
    a = oe.reflect(whole)
    b = oe.reflect(view)
    _same(a[0], b[0], 'view gb')
    _same(a[1], b[1], 'view lb')
    ref = scr.expose(a[0])
    out, _ = oe.reflect(view, exposeOn=scr)
    _same(out, ref, 'view fused')


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['dcm_si111', 'dcm_si111_asym'])
def test_double_reflect_expose_fused_equals_two_calls(name, cuda_device):
    from helpers import product_beamline, product_oe_kwargs
    from test_dcm import DCM_KEYS
    from xrt_b200 import oes, materials
    from xrt_b200.planes import Screen
    g = golden(name)
    si = materials.CrystalSi(hkl=(1, 1, 1), tK=float(g['tK']))
    kw = product_oe_kwargs(g)
    kw.update({k: float(g[k]) for k in DCM_KEYS})
    dcm = oes.DCM(product_beamline(g), 'dcm', material=si, material2=si, alpha=float(g['alpha']) or None,
                  limPhysX2=list(g['surfPhysX2']), limPhysY2=list(g['surfPhysY2']), **kw)
    dcm.lostNum = int(g['lostNum'])
    scr = Screen(dcm.bl, 'scr', center=[0., float(g['center'][1]) + 1e4, 20.])
    beam = device_beam(oracle_beam(g, 'in_'))
    gb, lo1, lo2 = dcm.double_reflect(beam)
    ref = scr.expose(gb)
    out, l1, l2 = dcm.double_reflect(beam, exposeOn=scr)
    _same(out, ref, name)
    _same(l1, lo1, name + ' lo1')
    _same(l2, lo2, name + ' lo2')
