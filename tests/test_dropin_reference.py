"""Drop-in at the reference's own call sites (SURVEY 7 'minimum slice').

tools/gen_dropin_calls.py ran the UNMODIFIED reference -- `rw.diffract()` with `waves.waveCL` set and
`Undulator.build_I_map()` with `Undulator.ucl` set (waves.py:493-498, :854-896; sybase.py:1094, synchr.py:2110-2193) --
with a recorder at the `run_parallel` boundary and stored the exact arguments those call sites produce together with
the reference's numpy results for them (and checked that the reference's OpenCL branch, fed with those results, ends
where its numpy branch ends).  Here the recorded calls are replayed, byte for byte, through XRT_CUDA.run_parallel.
(The reference itself cannot be imported on the GPU box, and there is no GPU in the build container.)"""
import numpy as np
import pytest

from conftest import golden
from oracle import kirchhoff as ok


def _call(g, prefix):
    kw = dict(kernelName=str(g[prefix + 'kernelName']), dimension=int(g[prefix + 'dimension']))
    for group in ('scalarArgs', 'slicedROArgs', 'nonSlicedROArgs', 'slicedRWArgs'):
        lst = []
        for i in range(int(g[prefix + group + '_n'])):
            a = g['%s%s_%d' % (prefix, group, i)]
            if bool(g['%s%s_%d_forder' % (prefix, group, i)]):
                a = np.asfortranarray(a)
            elif a.ndim == 0:
                a = a[()]
            lst.append(a)
        kw[group] = lst
    kw['nonSlicedRWArgs'] = None
    return kw


def test_recorded_calls_have_the_reference_layout():
    """What the reference's call sites pass (no GPU needed): kernel names, argument counts and dtypes, the (4, N)
    Fortran-ordered float3 buffers of waves.py:874-881 -- and XRT_CUDA's marshalling accepts them as they are."""
    from xrt_b200.cuda_backend import _as_vec_buffer
    g = golden('dropin_calls')
    k = _call(g, 'kir_')
    assert k['kernelName'] == 'integrate_kirchhoff'
    N = int(k['scalarArgs'][0])
    assert len(k['slicedROArgs']) == 3 and len(k['nonSlicedROArgs']) == 6 and len(k['slicedRWArgs']) == 5
    pos, nrm = k['nonSlicedROArgs'][4], k['nonSlicedROArgs'][5]
    assert pos.shape == (4, N) and pos.flags['F_CONTIGUOUS'] and nrm.shape == (4, N)
    assert all(a.dtype == np.complex128 for a in k['slicedRWArgs'])
    buf, stride = _as_vec_buffer(pos, N)
    assert stride == 4 and buf is pos  # taken as it is: N records of 4 doubles
    u = _call(g, 'und_')
    assert u['kernelName'] == 'undulator' and len(u['scalarArgs']) == 4
    assert len(u['slicedROArgs']) == 6 and len(u['nonSlicedROArgs']) == 6 and len(u['slicedRWArgs']) == 2
    assert int(u['scalarArgs'][3]) == len(u['nonSlicedROArgs'][0])


@pytest.mark.gpu
def test_replay_reference_diffract_call(cuda_device):
    from xrt_b200.cuda_backend import XRT_CUDA
    from xrt_b200.physconsts import CHBAR
    g = golden('dropin_calls')
    kw = _call(g, 'kir_')
    waveCL = XRT_CUDA(r'diffract.cl')  # what waves.py:493 constructs
    assert waveCL.lastTargetOpenCL is not None and waveCL.cl_precisionF is np.float64
    res = waveCL.run_parallel(**kw)
    assert all(a is b for a, b in zip(res, kw['slicedRWArgs']))
    xM, yM, zM = kw['slicedROArgs']
    nl, Es, Ep, k, pos, nrm = kw['nonSlicedROArgs']
    n3 = [nrm[i] for i in range(3)]
    sc = ok.term_magnitude_sum(pos[0], pos[1], pos[2], n3, nl, Es, Ep, k * CHBAR / 1e7, xM, yM, zM)
    for r, name, s in zip(res, ('Es', 'Ep', 'aE', 'bE', 'cE'), sc):
        # k was formed by the reference (E / CHBAR * 1e7) and travels as it is; the scale goes through E again, which
        # only moves the tolerance by an ulp
        assert (np.abs(r - g['kir_numpy_' + name]) / s).max() < 1e-13, name
    # OpenCL sign / scale convention of cl/diffract.cl:137-149 on request: constant factors that waves.diffract
    # normalises away (last paragraph)
    waveCL.kirchhoffConvention = 'opencl'
    kw2 = _call(g, 'kir_')
    res2 = waveCL.run_parallel(**kw2)
    np.testing.assert_allclose(res2[0], -g['kir_numpy_Es'], rtol=0, atol=1e-13 * sc[0].max())


@pytest.mark.gpu
def test_replay_reference_undulator_call(cuda_device):
    from xrt_b200.cuda_backend import XRT_CUDA
    g = golden('dropin_calls')
    kw = _call(g, 'und_')
    ucl = XRT_CUDA(r'undulator.cl')  # what IntegratedSource._set_cl constructs (sybase.py:1094)
    Is, Ip = ucl.run_parallel(**kw)
    # scale: sum over the nodes of |weight| (ag) times the largest integrand is what the raw sums are made of; the
    # golden-based test (test_undulator.py) bounds the error per term -- here a uniform 1e-12 of the largest sum
    ref_s, ref_p = g['und_numpy_Is'], g['und_numpy_Ip']
    scale = max(np.abs(ref_s).max(), np.abs(ref_p).max())
    assert np.abs(Is - ref_s).max() < 1e-11 * scale
    assert np.abs(Ip - ref_p).max() < 1e-11 * scale
    # and the reference's own post-processing of the kernel's output (synchr.py:2176-2193: a real factor per sample
    # times Is) lands on its numpy result
    fac = g['und_Es'] / np.where(ref_s == 0, 1, ref_s)
    assert np.abs(fac * Is - g['und_Es']).max() < 1e-11 * np.abs(g['und_Es']).max()
