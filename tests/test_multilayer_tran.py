"""f2: transmitted multilayers -- Multilayer(geom='transmitted') (materials/multilayer.py:388-516; the reference's
kernel get_amplitude_graded_multilayer_tran, cl/materials.cl) stand-alone and inside OE.reflect (toWhere = 2: the rays
pass straight, reflect.py:749-757); cf. the reference's tests/raycing/test_multilayer_transmission.py.  Oracle bit-exact
on the reference's outputs (CPU); CUDA against the same goldens (GPU)."""
import numpy as np
import pytest

from conftest import golden
from helpers import (oracle_spec, oracle_beam, product_beamline, product_oe_kwargs, device_beam, compare_beams,
                     BEAM_FIELDS)
from oracle import materials as om, reflect as orf

OUT_FIELDS = BEAM_FIELDS + ('theta',)
KW = dict(tThickness=15., bThickness=25., nPairs=30, substThickness=100., idThickness=1.5, geom='transmitted')


def _oracle_ml():
    return om.Multilayer(om.Material('Cr', rho=7.18), KW['tThickness'], om.Material('C', rho=2.26), KW['bThickness'],
                         KW['nPairs'], om.Material('C', rho=2.26), idThickness=KW['idThickness'],
                         substThickness=KW['substThickness'], geom='transmitted')


def test_oracle_multilayer_transmission_bit_exact():
    g = golden('multilayer_tran_amplitudes')
    ts, tp = _oracle_ml().get_amplitude(g['E'].copy(), g['bdn'].copy())
    assert np.array_equal(ts, g['ts']) and np.array_equal(tp, g['tp'])
    assert 0.01 < (np.abs(ts)**2).max() < 1.0


def test_oracle_multilayer_transmission_reflect_bit_exact():
    g = golden('reflect_multilayer_tran')
    gb, lb = orf.reflect(oracle_spec(g, 'flat', _oracle_ml()), oracle_beam(g, 'in_'))
    for pre, b in (('gb_', gb), ('lb_', lb)):
        for f in OUT_FIELDS:
            if pre + f in g.files:
                assert np.array_equal(getattr(b, f), g[pre + f]), pre + f
    good = g['gb_state'] == 1
    assert np.array_equal(g['gb_a'][good], g['in_a'][good])  # the rays pass straight


def _product_ml():
    from xrt_b200 import materials
    return materials.Multilayer(tLayer=materials.Material('Cr', rho=7.18), bLayer=materials.Material('C', rho=2.26),
                                substrate=materials.Material('C', rho=2.26), **KW)


@pytest.mark.gpu
def test_cuda_multilayer_transmission_amplitude(cuda_device):
    g = golden('multilayer_tran_amplitudes')
    ts, tp = _product_ml().get_amplitude(g['E'], g['bdn'])
    assert np.abs(ts - g['ts']).max() < 1e-12, np.abs(ts - g['ts']).max()  # |t| <= 0.27 here
    assert np.abs(tp - g['tp']).max() < 1e-12, np.abs(tp - g['tp']).max()


@pytest.mark.gpu
def test_cuda_multilayer_transmission_reflect(cuda_device):
    from xrt_b200 import oes
    g = golden('reflect_multilayer_tran')
    oe = oes.OE(product_beamline(g), 'mlt', material=_product_ml(), **product_oe_kwargs(g))
    oe.lostNum = int(g['lostNum'])
    gb, lb = oe.reflect(device_beam(oracle_beam(g, 'in_')))
    bad = compare_beams(gb.to_numpy(), {f: g['gb_' + f] for f in OUT_FIELDS if 'gb_' + f in g.files}, OUT_FIELDS,
                        1e-12, 'gb_')
    bad += compare_beams(lb.to_numpy(), {f: g['lb_' + f] for f in OUT_FIELDS if 'lb_' + f in g.files}, OUT_FIELDS,
                         1e-12, 'lb_')
    assert not bad, bad
    assert np.array_equal(gb.to_numpy()['state'], g['gb_state'])


@pytest.mark.gpu
def test_cuda_run_parallel_multilayer_tran_kernel(cuda_device):
    """Kernel-dispatch level: XRT_CUDA.run_parallel('get_amplitude_graded_multilayer_tran', ...) with the argument
    lists Multilayer.get_amplitude hands to its accelerator (multilayer.py:485-503), fed with the oracle's interface
    amplitudes: the recursion alone reproduces the reference's (ts, tp)."""
    from xrt_b200.cuda_backend import XRT_CUDA
    g = golden('multilayer_tran_amplitudes')
    ml = _oracle_ml()
    f = ml.fresnel_terms(g['E'].copy(), g['bdn'].copy())
    order = ['rvt', 'tvt', 'rbs', 'tbs', 'rsv', 'tsv', 'rbt', 'tbt', 'rtb', 'ttb']
    f['rbt_s'], f['rbt_p'] = -f['rtb_s'], -f['rtb_p']
    sliced = [np.ascontiguousarray(f[k + sfx]) for k in order for sfx in ('_s', '_p')] + \
        [np.ascontiguousarray(f[k]) for k in ('Qt', 'Qb', 'Qs')]
    n = len(g['E'])
    ti_s, ti_p = np.zeros(n, dtype=complex), np.zeros(n, dtype=complex)
    ucl = XRT_CUDA(targetOpenCL='CUDA', precisionOpenCL='float64')
    out = ucl.run_parallel('get_amplitude_graded_multilayer_tran',
                           [np.int32(ml.nPairs), np.float64(ml.substThickness)], sliced,
                           [np.float64(ml.dti), np.float64(ml.dbi)], [ti_s, ti_p], None, n)
    assert out[0] is ti_s and out[1] is ti_p
    assert np.abs(ti_s - g['ts']).max() < 1e-13 and np.abs(ti_p - g['tp']).max() < 1e-13
