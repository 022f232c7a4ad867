"""Known-answer tests of the device math helpers (common.cuh) the hot kernels are built from, against numpy /
mpmath-free references: the reference's phases reach 5e11 rad (Kirchhoff) and 1e12 rad (exp(i k t) in reflect), where
CUDA's own sincos leaves its fast path; r = sqrt(.) must be the correctly rounded value numpy produces."""
import numpy as np
import pytest


def _probe(which, x):
    from xrt_b200 import _lib
    lib = _lib.load()
    _lib.init()
    x = np.ascontiguousarray(x, dtype=np.float64)
    y0 = np.empty_like(x)
    y1 = np.empty_like(x)
    _lib.check(lib.xrt_b200_math_probe_host(which, len(x), _lib.hptr(x), _lib.hptr(y0), _lib.hptr(y1)), 'math_probe')
    return y0, y1


def _exact_sincos(x):
    """sin / cos of float64 arguments up to 2^50 with numpy's longdouble reduction (x87 80-bit: 64-bit mantissa
    is not enough by itself for 1e12 rad, so reduce with a two-term pi first)."""
    from fractions import Fraction
    import math
    # exact reduction with integers: pi as a 200-bit rational
    PI = Fraction(314159265358979323846264338327950288419716939937510582097494459230781640628620899, 10**80)
    out_s = np.empty(len(x))
    out_c = np.empty(len(x))
    for i, v in enumerate(x):
        f = Fraction(float(v))
        n = round(f / (PI / 2))
        r = float(f - n * (PI / 2))
        s, c = math.sin(r), math.cos(r)
        q = n % 4
        out_s[i], out_c[i] = [(s, c), (c, -s), (-s, -c), (-c, s)][q]
    return out_s, out_c


@pytest.mark.gpu
def test_sincos_helpers_on_huge_phases(cuda_device):
    rng = np.random.RandomState(1)
    x = np.concatenate([rng.uniform(-1e12, 1e12, 3000), rng.uniform(-100, 100, 2000), rng.uniform(-2**49, 2**49, 1000),
                        np.array([0., 1e-300, np.pi / 2, np.pi, 1.5707963267948966, 3.141592653589793e11])])
    s_ref, c_ref = _exact_sincos(x)
    s, c = _probe(0, x)
    assert np.abs(s - s_ref).max() < 4e-16 and np.abs(c - c_ref).max() < 4e-16
    s1, _ = _probe(1, x)  # sin_big (mod pi, one odd polynomial)
    assert np.abs(s1 - s_ref).max() < 4e-16
    # sincos_tab (mod pi/2048, 4096-entry table + two-term polynomials), the Kirchhoff kernel's: valid below 2^40 pi rad
    from xrt_b200 import kernels
    m = np.abs(x) < kernels.PHASE_LIMIT_TABLE_SINCOS
    assert m.sum() > 5000  # the 1e12-rad phases are inside the range
    s2, c2 = _probe(8, x[m])
    assert np.abs(s2 - s_ref[m]).max() < 5e-16 and np.abs(c2 - c_ref[m]).max() < 5e-16


@pytest.mark.gpu
def test_reciprocal_and_sqrt_helpers(cuda_device):
    rng = np.random.RandomState(2)
    x = np.concatenate([rng.uniform(1e-9, 1e-7, 4000), 10.**rng.uniform(-30, 30, 4000), rng.uniform(0.5, 2, 4000)])
    for which, ulps in ((2, 1.0), (6, 1.5)):
        y, _ = _probe(which, x)
        assert (np.abs(y * x - 1).max()) < 4e-16
        assert (np.abs(y - 1 / x) / np.spacing(1 / x)).max() <= ulps, which
    r, inv = _probe(3, x)
    assert np.array_equal(r, np.sqrt(x))  # correctly rounded, like numpy's
    assert (np.abs(inv - 1 / np.sqrt(x)) / np.spacing(1 / np.sqrt(x))).max() <= 2
    r, hinv = _probe(7, x)
    assert np.array_equal(r, np.sqrt(x))
    assert (np.abs(hinv - 0.5 / np.sqrt(x)) / np.spacing(0.5 / np.sqrt(x))).max() <= 2


@pytest.mark.gpu
def test_bessel_k_fractional_orders(cuda_device):
    from scipy import special
    x = np.concatenate([np.geomspace(1e-8, 2, 3000), np.linspace(2.0000001, 60, 3000), np.geomspace(60, 700, 500)])
    k13, k23 = _probe(4, x)
    r13, r23 = special.kv(1/3., x), special.kv(2/3., x)
    ok = r13 > 1e-300  # the last points underflow
    assert (np.abs(k13 - r13)[ok] / r13[ok]).max() < 1.5e-13
    assert (np.abs(k23 - r23)[ok] / r23[ok]).max() < 1.5e-13
    assert np.all(k13[~ok] < 1e-299)
    k13, k23 = _probe(4, np.array([0., 900., np.inf]))
    assert np.isinf(k13[0]) and k13[1] == 0 and k13[2] == 0 and k23[2] == 0


@pytest.mark.gpu
def test_sin_table_of_the_undulator_pair_loop(cuda_device):
    """sin_tab (1024-entry table + r - r^3/6 + r^5/120): the node-pair step's sin, phases O(100) rad."""
    rng = np.random.RandomState(5)
    x = np.concatenate([rng.uniform(-300, 300, 6000), rng.uniform(-1e6, 1e6, 2000),
                        np.array([0., np.pi / 512, np.pi / 1024, -np.pi, 1e-300, 100.0])])
    s_ref, _ = _exact_sincos(x)
    s, _ = _probe(9, x)
    assert np.abs(s - s_ref).max() < 4e-16


@pytest.mark.gpu
def test_unchecked_division_and_sqrt_are_correctly_rounded(cuda_device):
    """div_nr / sqrt_nr (common.cuh): nvcc's fast path of a / b and sqrt(a) without the range test and the slow-path
    call -- for ordinary operands the results must be the IEEE ones, bit for bit (reflect.cu relies on it: the toroid
    surface, the secant step, the Fresnel and Bragg amplitudes)."""
    rng = np.random.RandomState(6)
    x = np.concatenate([rng.uniform(-2, 2, 200000), 10.**rng.uniform(-200, 200, 200000) * rng.choice([-1, 1], 200000),
                        rng.uniform(0.999999, 1.000001, 100000), np.array([1.0, -1.0, 3.0, 1e-290, 1e290, 0.1])])
    x = x[x != 0]
    inv, rt = _probe(10, x)
    assert np.array_equal(inv, 1.0 / x)
    assert np.array_equal(rt, np.sqrt(np.abs(x)))
    a = rng.uniform(-1e3, 1e3, 400000) * 10.**rng.uniform(-30, 30, 400000)
    b = rng.uniform(-1e3, 1e3, 400000) * 10.**rng.uniform(-30, 30, 400000)
    b[b == 0] = 1.0
    a[::1000] = 0.0  # zero dividends
    pairs = np.empty(2 * len(a))
    pairs[0::2] = a
    pairs[1::2] = b
    q, _ = _probe(11, pairs)
    with np.errstate(all='ignore'):
        assert np.array_equal(q[0::2], a / b)
        sel = a != 0
        assert np.array_equal(q[1::2][sel], (b / a)[sel])
