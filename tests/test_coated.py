"""a12 / f2: Coated (materials/multilayer.py:528-584) -- a reflective layer on a substrate with surface and substrate
roughness, kind 'mirror': Parratt recursion amplitude + specular reflection.  Oracle bit-exact on the reference's
outputs (CPU); the CUDA amplitude kernel and the fused reflect kernel against the same goldens (GPU)."""
import numpy as np
import pytest

from conftest import golden
from helpers import (oracle_spec, oracle_beam, product_beamline, product_oe_kwargs, device_beam, compare_beams,
                     BEAM_FIELDS)
from oracle import materials as om, reflect as orf

OUT_FIELDS = BEAM_FIELDS + ('theta',)


def _oracle_coated(g):
    return om.Coated(om.Material('Pt', rho=21.45), float(g['cThickness']), om.Material('Si', rho=2.33),
                     surfaceRoughness=float(g['surfaceRoughness']), substRoughness=float(g['substRoughness']))


def test_oracle_coated_amplitude_bit_exact():
    g = golden('coated_amplitudes')
    rs, rp = _oracle_coated(g).get_amplitude(g['E'].copy(), g['bdn'].copy(), True)
    assert np.array_equal(rs, g['rs']) and np.array_equal(rp, g['rp'])
    assert 0.9 < (np.abs(rs)**2).max() <= 1.0


def test_oracle_coated_reflect_bit_exact():
    g = golden('reflect_coated_toroid')
    sp = oracle_spec(g, 'toroid', _oracle_coated(g), surfParams=dict(R=float(g['R']), r=float(g['r'])))
    gb, lb = orf.reflect(sp, oracle_beam(g, 'in_'))
    for pre, b in (('gb_', gb), ('lb_', lb)):
        for f in OUT_FIELDS:
            if pre + f in g.files:
                assert np.array_equal(getattr(b, f), g[pre + f]), pre + f


def _product_coated(g):
    from xrt_b200 import materials
    return materials.Coated(coating=materials.Material('Pt', rho=21.45), cThickness=float(g['cThickness']),
                            surfaceRoughness=float(g['surfaceRoughness']),
                            substrate=materials.Material('Si', rho=2.33), substRoughness=float(g['substRoughness']))


@pytest.mark.gpu
def test_cuda_coated_amplitude(cuda_device):
    g = golden('coated_amplitudes')
    rs, rp = _product_coated(g).get_amplitude(g['E'], g['bdn'], True)
    assert np.abs(rs - g['rs']).max() < 1e-11, np.abs(rs - g['rs']).max()
    assert np.abs(rp - g['rp']).max() < 1e-11, np.abs(rp - g['rp']).max()


@pytest.mark.gpu
def test_cuda_coated_reflect(cuda_device):
    from xrt_b200 import oes
    g = golden('reflect_coated_toroid')
    oe = oes.ToroidMirror(product_beamline(g), 'tm', R=float(g['R']), r=float(g['r']), material=_product_coated(g),
                          **product_oe_kwargs(g))
    oe.lostNum = int(g['lostNum'])
    gb, lb = oe.reflect(device_beam(oracle_beam(g, 'in_')))
    bad = compare_beams(gb.to_numpy(), {f: g['gb_' + f] for f in OUT_FIELDS if 'gb_' + f in g.files}, OUT_FIELDS,
                        1e-11, 'gb_')
    bad += compare_beams(lb.to_numpy(), {f: g['lb_' + f] for f in OUT_FIELDS if 'lb_' + f in g.files}, OUT_FIELDS,
                         1e-11, 'lb_')
    assert not bad, bad
    assert np.array_equal(gb.to_numpy()['state'], g['gb_state'])
