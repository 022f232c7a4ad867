"""f2: finite-thickness crystal amplitudes -- Bragg / Laue, reflected / transmitted (materials/crystal.py:598-616) --
stand-alone and inside OE.reflect (Laue deflection sign reflect.py:608, 'transmitted' = pass straight :756).
Oracle bit-exact on the reference's outputs (CPU); CUDA against the same goldens (GPU)."""
import numpy as np
import pytest

from conftest import golden
from helpers import (oracle_spec, oracle_beam, product_beamline, product_oe_kwargs, device_beam, compare_beams,
                     BEAM_FIELDS)
from oracle import materials as om, reflect as orf

THIN = {'bragg_thin': ('Bragg reflected', 0.005), 'bragg_t': ('Bragg transmitted', 0.05),
        'laue_r': ('Laue reflected', 0.1), 'laue_t': ('Laue transmitted', 0.1)}
OUT_FIELDS = BEAM_FIELDS + ('theta',)


@pytest.mark.parametrize('key', sorted(THIN))
def test_oracle_thin_crystal_amplitude_bit_exact(key):
    g = golden('amplitude_crystals_thin')
    geom, t = THIN[key]
    si = om.CrystalDiamond((1, 1, 1), float(g[key + '_d']), 'Si', t=t, geom=geom)
    cs, cp = si.get_amplitude(g[key + '_E'], g[key + '_g0'].copy(), g[key + '_gh'].copy(), g[key + '_hns0'].copy())
    assert np.array_equal(cs, g[key + '_rs']) and np.array_equal(cp, g[key + '_rp'])
    assert 0.1 < (np.abs(cs)**2).max() <= 1.0


@pytest.mark.parametrize('key', sorted(THIN))
def test_oracle_thin_crystal_reflect_bit_exact(key):
    g = golden('reflect_crystal_' + key)
    geom, t = THIN[key]
    si = om.CrystalDiamond((1, 1, 1), float(g['d']), 'Si', t=t, geom=geom)
    sp = oracle_spec(g, 'flat', si, alpha=float(g['alpha']) or None)
    gb, lb = orf.reflect(sp, oracle_beam(g, 'in_'))
    for pre, b in (('gb_', gb), ('lb_', lb)):
        for f in OUT_FIELDS:
            if pre + f in g.files:
                assert np.array_equal(getattr(b, f), g[pre + f]), pre + f


ATOL = 1e-11


@pytest.mark.gpu
@pytest.mark.parametrize('key', sorted(THIN))
def test_cuda_thin_crystal_amplitude(key, cuda_device):
    from xrt_b200.materials import CrystalSi
    g = golden('amplitude_crystals_thin')
    geom, t = THIN[key]
    si = CrystalSi(hkl=(1, 1, 1), geom=geom, t=t)
    assert si.d == float(g[key + '_d'])
    cs, cp = si.get_amplitude(g[key + '_E'], g[key + '_g0'], g[key + '_gh'], g[key + '_hns0'])
    assert np.abs(cs - g[key + '_rs']).max() < ATOL, np.abs(cs - g[key + '_rs']).max()
    assert np.abs(cp - g[key + '_rp']).max() < ATOL, np.abs(cp - g[key + '_rp']).max()


@pytest.mark.gpu
@pytest.mark.parametrize('key', sorted(THIN))
def test_cuda_thin_crystal_reflect(key, cuda_device):
    from xrt_b200 import oes
    from xrt_b200.materials import CrystalSi
    g = golden('reflect_crystal_' + key)
    geom, t = THIN[key]
    si = CrystalSi(hkl=(1, 1, 1), geom=geom, t=t, tK=float(g['tK']))
    oe = oes.OE(product_beamline(g), 'xtal', material=si, alpha=float(g['alpha']) or None, **product_oe_kwargs(g))
    gb, lb = oe.reflect(device_beam(oracle_beam(g, 'in_')))
    bad = compare_beams(gb.to_numpy(), {f: g['gb_' + f] for f in OUT_FIELDS if 'gb_' + f in g.files}, OUT_FIELDS,
                        1e-11, 'gb_')
    bad += compare_beams(lb.to_numpy(), {f: g['lb_' + f] for f in OUT_FIELDS if 'lb_' + f in g.files}, OUT_FIELDS,
                         1e-11, 'lb_')
    assert not bad, bad
