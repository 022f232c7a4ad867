"""a12/a13: mirror Fresnel amplitudes and thick-Bragg crystal amplitudes -- oracle bit-exact on the reference's
outputs (CPU); CUDA kernels against the same goldens (GPU).  The crystal grid is the recipe of the reference's
rocking-curve comparison (tests/raycing/test_materials.py:239-352)."""
import numpy as np
import pytest

from conftest import golden
from oracle import materials as om

MIRRORS = {'Au': dict(elements='Au', rho=19.3), 'Pt': dict(elements='Pt', rho=21.45),
           'SiO2': dict(elements=('Si', 'O'), quantities=(1, 2), rho=2.2),
           'AuThin': dict(elements='Au', rho=19.3, kind='thin mirror', t=2e-5)}


def _crystal_keys(g):
    return sorted(set(k.rsplit('_', 1)[0] for k in g.files))


@pytest.mark.parametrize('name', sorted(MIRRORS))
def test_oracle_mirror_bit_exact(name):
    g = golden('amplitude_mirrors')
    res = om.Material(**MIRRORS[name]).get_amplitude(g['E'], g['beamInDotNormal'].copy())
    assert np.array_equal(res[0], g[name + '_rs']) and np.array_equal(res[1], g[name + '_rp'])


def test_oracle_crystal_bit_exact():
    g = golden('amplitude_crystals')
    for key in _crystal_keys(g):
        hkl = tuple(int(c) for c in key[2:5])
        si = om.CrystalDiamond(hkl, float(g[key + '_d']), 'Si')
        cs, cp = si.get_amplitude(g[key + '_E'], g[key + '_g0'].copy(), g[key + '_gh'].copy(),
                                  g[key + '_hns0'].copy())
        assert np.array_equal(cs, g[key + '_rs']) and np.array_equal(cp, g[key + '_rp']), key


def test_product_crystal_constants_match_reference():
    """d spacing of CrystalSi (thermal expansion) equals the reference's (crystals_basic.py:83-142)."""
    from xrt_b200.materials import CrystalSi
    g = golden('amplitude_crystals')
    assert CrystalSi(hkl=(1, 1, 1)).d == float(g['Si111_a+0_d'])
    assert CrystalSi(hkl=(3, 3, 3)).d == float(g['Si333_a+0_d'])
    gd = golden('dcm_si111')
    assert CrystalSi(hkl=(1, 1, 1), tK=float(gd['tK'])).d == float(gd['d'])


# |r| <= 1; differences come from libm-vs-CUDA transcendental functions amplified by the grazing-incidence /
# Bragg-peak cancellations that the reference carries itself (~1e4): assert 1e-11 absolute on amplitudes <= 1.
ATOL = 1e-11


@pytest.mark.gpu
@pytest.mark.parametrize('name', sorted(MIRRORS))
def test_cuda_mirror_amplitude(name, cuda_device):
    from xrt_b200.materials import Material
    g = golden('amplitude_mirrors')
    rs, rp = Material(**MIRRORS[name]).get_amplitude(g['E'], g['beamInDotNormal'])
    assert np.abs(rs - g[name + '_rs']).max() < ATOL
    assert np.abs(rp - g[name + '_rp']).max() < ATOL


@pytest.mark.gpu
def test_cuda_crystal_amplitude(cuda_device):
    from xrt_b200.materials import CrystalSi
    g = golden('amplitude_crystals')
    for key in _crystal_keys(g):
        hkl = tuple(int(c) for c in key[2:5])
        si = CrystalSi(hkl=hkl)
        cs, cp = si.get_amplitude(g[key + '_E'], g[key + '_g0'], g[key + '_gh'], g[key + '_hns0'])
        assert np.abs(cs - g[key + '_rs']).max() < ATOL, key
        assert np.abs(cp - g[key + '_rp']).max() < ATOL, key


@pytest.mark.gpu
def test_cuda_amplitude_out_of_table_raises(cuda_device):
    from xrt_b200.materials import Material
    with pytest.raises(ValueError):
        Material('Au', rho=19.3).get_amplitude(np.array([1e6]), np.array([-0.004]))


def test_host_alignment_helpers_match_oracle():
    """Crystal.get_Bragg_angle / get_dtheta (crystal.py:1105-1173): host-side numpy helpers of the product."""
    from xrt_b200 import materials
    si = materials.CrystalSi(hkl=(1, 1, 1))
    o = om.CrystalDiamond((1, 1, 1), si.d, 'Si')
    E = np.array([9000., 7000., 12000.])
    assert np.array_equal(si.get_Bragg_angle(E), o.get_Bragg_angle(E))
    assert np.array_equal(si.get_dtheta(E), o.get_dtheta(E))
    assert np.array_equal(si.get_dtheta(E, 0.05), o.get_dtheta(E, 0.05))


@pytest.mark.gpu
def test_run_parallel_get_amplitude_kernel_interface(cuda_device):
    """XRT_CUDA.run_parallel('get_amplitude', ...) with the argument list of the reference kernel
    (cl/materials.cl:715-748): the silicon cell given atom by atom (8 atoms of the diamond cubic cell, occupancy 1),
    scattering-factor tables as maxEl x 300 vectors; results against the reference's numpy Crystal.get_amplitude."""
    from xrt_b200.cuda_backend import XRT_CUDA
    from xrt_b200.materials import CrystalSi
    g = golden('amplitude_crystals')
    fcc = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0]])
    xyz = np.concatenate([fcc, fcc + 0.25])
    cl = XRT_CUDA(r'materials.cl')
    for key in _crystal_keys(g):
        hkl = tuple(int(c) for c in key[2:5])
        si = CrystalSi(hkl=hkl)
        el = si.elements[0]
        maxEl = len(xyz)
        nt = len(el.E)
        elements = np.zeros((maxEl, 8))
        elements[:, :3] = xyz
        elements[:, 4] = 1.
        elements[:, 5] = el.Z
        elements[:, 7] = nt
        f0cfs = np.tile(np.asarray(el.f0coeffs, dtype=float), maxEl)
        vec = lambda a: np.concatenate([np.pad(np.asarray(a, dtype=float), (0, 300 - nt))] * maxEl)  # noqa: E731
        E = g[key + '_E']
        n = len(E)
        rs = np.zeros(n, dtype=np.complex128)
        rp = np.zeros(n, dtype=np.complex128)
        out = cl.run_parallel(
            'get_amplitude', scalarArgs=[np.array(list(hkl) + [0], dtype=np.int32), np.float64(1.), np.float64(0.),
                                         np.int32(2), np.int32(maxEl)],
            slicedROArgs=[g[key + '_g0'], g[key + '_gh'], g[key + '_hns0'], E, np.full(n, si.d),
                          np.full(n, si.chiToF)],
            nonSlicedROArgs=[elements.ravel(), f0cfs, vec(el.E), vec(el.f1), vec(el.f2)],
            slicedRWArgs=[rs, rp], dimension=n)
        assert out[0] is rs and out[1] is rp  # RW arrays are written in place and returned (myopencl.py:550-583)
        assert np.abs(rs - g[key + '_rs']).max() < ATOL, key
        assert np.abs(rp - g[key + '_rp']).max() < ATOL, key
    # a cell of two different elements is refused, not silently approximated
    elements[1, 5] += 1
    with pytest.raises(NotImplementedError):
        cl.run_parallel('get_amplitude', scalarArgs=[np.array([1, 1, 1, 0]), 1., 0., 2, maxEl],
                        slicedROArgs=[g[key + '_g0'], g[key + '_gh'], g[key + '_hns0'], E, np.full(n, si.d),
                                      np.full(n, si.chiToF)],
                        nonSlicedROArgs=[elements.ravel(), f0cfs, vec(el.E), vec(el.f1), vec(el.f2)],
                        slicedRWArgs=[rs, rp], dimension=n)
