"""f2: Multilayer.get_amplitude (geom 'reflected', periodic and depth graded) and a multilayer mirror in OE.reflect
(materials/multilayer.py:254-508; oes/reflect.py:568-612, :999-1003) -- oracle bit-exact on the reference's
outputs (CPU); CUDA against the same goldens through the three entry points the reference has: the material's
get_amplitude, the accelerator call run_parallel('get_amplitude_graded_multilayer') and OE.reflect (GPU)."""
import numpy as np
import pytest

from conftest import golden
from helpers import (oracle_spec, oracle_beam, product_beamline, product_oe_kwargs, device_beam, compare_beams,
                     BEAM_FIELDS)
from oracle import materials as om, reflect as orf

STACKS = {
    'wsi': dict(t=('W', 19.3), tThickness=12., b=('Si', 2.33), bThickness=22., nPairs=60, s=('Si', 2.33),
                idThickness=2.),
    'moc_graded': dict(t=('Mo', 10.28), tThickness=25., b=('C', 2.26), bThickness=35., nPairs=40, s=('Si', 2.33),
                       tThicknessLow=15., bThicknessLow=20., idThickness=3., power=2.),
}
OUT_FIELDS = BEAM_FIELDS + ('theta',)


def _stack(mod, name):
    kw = dict(STACKS[name])
    mats = {k: mod.Material(kw[k][0], rho=kw[k][1]) for k in ('t', 'b', 's')}
    for k in ('t', 'b', 's'):
        kw.pop(k)
    return mod.Multilayer(mats['t'], kw.pop('tThickness'), mats['b'], kw.pop('bThickness'), kw.pop('nPairs'),
                          mats['s'], **kw)


@pytest.mark.parametrize('name', sorted(STACKS))
def test_oracle_multilayer_amplitude_bit_exact(name):
    g = golden('multilayer_amplitudes')
    ml = _stack(om, name)
    assert np.array_equal(ml.dti, g[name + '_dti']) and np.array_equal(ml.dbi, g[name + '_dbi'])
    rs, rp = ml.get_amplitude(g[name + '_E'], g[name + '_bdn'])
    assert np.array_equal(rs, g[name + '_rs']) and np.array_equal(rp, g[name + '_rp'])
    assert (np.abs(rs)**2).max() > 0.8  # the scan crosses the first Bragg peak


def test_oracle_multilayer_mirror_bit_exact():
    g = golden('reflect_multilayer_flat')
    sp = oracle_spec(g, 'flat', _stack(om, 'wsi'))
    info = {}
    gb, lb = orf.reflect(sp, oracle_beam(g, 'in_'), info=info)
    for pre, b in (('gb_', gb), ('lb_', lb)):
        for f in OUT_FIELDS:
            if pre + f in g.files:
                assert np.array_equal(getattr(b, f), g[pre + f]), pre + f
    assert np.array_equal(info['footprint'], g['footprint'])


def test_product_thickness_laws_match_reference():
    from xrt_b200 import materials
    g = golden('multilayer_amplitudes')
    for name in STACKS:
        ml = _stack(materials, name)
        assert np.array_equal(ml.dti, g[name + '_dti']) and np.array_equal(ml.dbi, g[name + '_dbi']), name
        assert ml.d == float(g[name + '_d'])
        thB = float(ml.get_Bragg_angle(9000.) - ml.get_dtheta(9000.))
        assert abs(thB - float(g[name + '_thB'])) <= 1e-15, name


# |r| <= 1; the recursion multiplies 2 x nPairs phase factors whose exp/sincos differ by <= 1 ulp between libm
# and CUDA: 1e-11 absolute
ATOL = 1e-11


@pytest.mark.gpu
@pytest.mark.parametrize('name', sorted(STACKS))
def test_cuda_multilayer_amplitude(name, cuda_device):
    from xrt_b200 import materials
    g = golden('multilayer_amplitudes')
    rs, rp = _stack(materials, name).get_amplitude(g[name + '_E'], g[name + '_bdn'])
    assert np.abs(rs - g[name + '_rs']).max() < ATOL
    assert np.abs(rp - g[name + '_rp']).max() < ATOL


@pytest.mark.gpu
@pytest.mark.parametrize('name', sorted(STACKS))
def test_cuda_run_parallel_graded_multilayer(name, cuda_device):
    """The reference's own accelerator call site (multilayer.py:471-484) served by XRT_CUDA.run_parallel."""
    from xrt_b200.cuda_backend import XRT_CUDA
    g = golden('multilayer_amplitudes')
    ml = _stack(om, name)
    f = ml.fresnel_terms(g[name + '_E'], g[name + '_bdn'])
    ri_s, ri_p = np.zeros_like(f['rbs_s']), np.zeros_like(f['rbs_s'])
    ucl = XRT_CUDA('materials.cl', targetOpenCL='CUDA')
    out = ucl.run_parallel('get_amplitude_graded_multilayer', [np.int32(ml.nPairs)],
                           [f[k] for k in ('rbs_s', 'rbs_p', 'rtb_s', 'rtb_p', 'rvt_s', 'rvt_p', 'Qt', 'Qb')],
                           [np.float64(ml.dti), np.float64(ml.dbi)], [ri_s, ri_p], None, len(ri_s))
    assert out[0] is ri_s and out[1] is ri_p
    assert np.abs(ri_s - g[name + '_rs']).max() < ATOL
    assert np.abs(ri_p - g[name + '_rp']).max() < ATOL


@pytest.mark.gpu
def test_cuda_multilayer_mirror_vs_golden(cuda_device):
    from xrt_b200 import oes, materials
    g = golden('reflect_multilayer_flat')
    oe = oes.OE(product_beamline(g), 'mlm', material=_stack(materials, 'wsi'), **product_oe_kwargs(g))
    gb, lb = oe.reflect(device_beam(oracle_beam(g, 'in_')), withFootprint=True)
    bad = compare_beams(gb.to_numpy(), {f: g['gb_' + f] for f in OUT_FIELDS if 'gb_' + f in g.files}, OUT_FIELDS,
                        1e-11, 'gb_')
    bad += compare_beams(lb.to_numpy(), {f: g['lb_' + f] for f in OUT_FIELDS if 'lb_' + f in g.files}, OUT_FIELDS,
                         1e-11, 'lb_')
    assert not bad, bad
    assert np.abs(oe.footprint[-1] - g['footprint']).max() <= 1e-12 * np.abs(g['footprint']).max()
