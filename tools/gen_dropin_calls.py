#!/usr/bin/env python
"""Record what the UNMODIFIED reference passes through its accelerator boundary, and what it gets back.

The reference's two live accelerator call sites on the hot path are `waves._diffraction_integral_CL` (waves.py:854-896,
kernel 'integrate_kirchhoff') and `Undulator._build_I_map_CL` (sources/synchr.py:2110-2193, kernel 'undulator'); both go
through `XRT_CL.run_parallel`.  pyopencl does not exist here and /root/reference does not exist on the GPU box, so the
reference cannot be run WITH our library in one process.  This script does the next closest thing, in this container:
it puts a recorder where `waves.waveCL` / `Undulator.ucl` stand (same attribute surface), runs the reference's own
`rw.diffract()` and `Undulator.build_I_map()` unmodified down their OpenCL branch, and stores

  * the exact `run_parallel` arguments they produced (names, order, dtypes, the (4, N) Fortran float3 buffers), and
  * the results of the reference's numpy implementation of the same kernel for those arguments
    (`_diffraction_integral_conv`, `Undulator._sp_sum`), which the recorder also hands back so that the reference's
    post-processing runs to the end; the final outputs of the OpenCL branch are compared here with the numpy branch.

tests/test_dropin_reference.py replays the recorded calls through XRT_CUDA.run_parallel on the GPU.

    python tools/gen_dropin_calls.py      ->  tests/golden/dropin_calls.npz
"""
import copy
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, '/root/reference')
sys.path.insert(0, os.path.join(HERE, 'ref_stubs'))

import numpy as np  # noqa: E402

import xrt.backends.raycing as raycing  # noqa: E402
import xrt.backends.raycing.sources as rs  # noqa: E402
import xrt.backends.raycing.screens as rsc  # noqa: E402
import xrt.backends.raycing.apertures as ra  # noqa: E402
import xrt.backends.raycing.waves as rw  # noqa: E402
from xrt.backends.raycing import physconsts as rpc  # noqa: E402

raycing._VERBOSITY_ = 0


class Recorder(object):
    """Stands where XRT_CL stands (myopencl.py:86-97): the attributes the call sites read, and run_parallel."""
    cl_precisionF = np.float64
    cl_precisionC = np.complex128
    cl_ctx = ['recorder']
    cl_queue = [None]
    cl_program = [None]
    cl_is_blocking = True
    lastTargetOpenCL = 'auto'

    def __init__(self, compute):
        self.calls = []
        self.compute = compute

    def run_parallel(self, kernelName='', scalarArgs=None, slicedROArgs=None, nonSlicedROArgs=None,
                     slicedRWArgs=None, nonSlicedRWArgs=None, dimension=0, complexity=0, signal=None):
        self.calls.append(dict(kernelName=kernelName, scalarArgs=copy.deepcopy(scalarArgs),
                               slicedROArgs=copy.deepcopy(slicedROArgs),
                               nonSlicedROArgs=copy.deepcopy(nonSlicedROArgs),
                               slicedRWArgs=copy.deepcopy(slicedRWArgs), nonSlicedRWArgs=nonSlicedRWArgs,
                               dimension=dimension))
        res = self.compute()
        for a, r in zip(slicedRWArgs, res):
            a[...] = r
        return tuple(slicedRWArgs)


def pack_call(prefix, call, out):
    out[prefix + 'kernelName'] = np.array(call['kernelName'])
    out[prefix + 'dimension'] = np.int64(call['dimension'])
    for group in ('scalarArgs', 'slicedROArgs', 'nonSlicedROArgs', 'slicedRWArgs'):
        lst = call[group] or []
        out[prefix + group + '_n'] = np.int64(len(lst))
        for i, a in enumerate(lst):
            a = np.asarray(a)
            out['%s%s_%d' % (prefix, group, i)] = a
            out['%s%s_%d_forder' % (prefix, group, i)] = np.bool_(a.ndim == 2 and a.flags['F_CONTIGUOUS'] and
                                                                  not a.flags['C_CONTIGUOUS'])


def record_diffract(out):
    bl = raycing.BeamLine(azimuth=0.1)
    d1, d2 = 1e4, 2e4
    sa, ca = np.sin(0.1), np.cos(0.1)
    slit = ra.RectangularAperture(bl, 'slit', center=(d1*sa, d1*ca, 0), kind=('left', 'right', 'bottom', 'top'),
                                  opening=[-0.05, 0.05, -0.04, 0.04])
    scr = rsc.Screen(bl, 'scr', center=(d2*sa, d2*ca, 0))
    N = 2500
    rng = np.random.RandomState(303)
    oeLocal = rs.Beam(N, withAmplitudes=True)
    oeLocal.x[:] = rng.uniform(-0.05, 0.05, N)
    oeLocal.z[:] = rng.uniform(-0.04, 0.04, N)
    oeLocal.y[:] = 0.
    r = np.sqrt(oeLocal.x**2 + d1**2 + oeLocal.z**2)
    oeLocal.a[:], oeLocal.b[:], oeLocal.c[:] = oeLocal.x / r, d1 / r, oeLocal.z / r
    oeLocal.E[:] = 9000.
    k = oeLocal.E / rpc.CHBAR * 1e7
    oeLocal.Es[:] = np.exp(1j * k * r) / r * 1e4
    oeLocal.Ep[:] = 0.05 * np.exp(1j * (k * r + 0.3)) / r * 1e4
    oeLocal.Jss[:] = np.abs(oeLocal.Es)**2
    oeLocal.Jpp[:] = np.abs(oeLocal.Ep)**2
    oeLocal.Jsp[:] = oeLocal.Es * np.conj(oeLocal.Ep)
    oeLocal.state[:] = 1
    oeLocal.state[rng.rand(N) < 0.03] = 2
    xm = np.linspace(-0.4, 0.4, 40)
    zm = np.linspace(-0.3, 0.3, 33)
    # numpy branch (waveCL is None here: pyopencl is absent), with the integral's own result captured
    conv = {}
    orig = rw._diffraction_integral_conv

    def spy(*a, **kw):
        res = orig(*a, **kw)
        conv['res'] = [np.array(v) for v in res]
        return res
    rw._diffraction_integral_conv = spy
    wave_np = scr.prepare_wave(slit, xm, zm)
    oe1 = copy.deepcopy(oeLocal)
    glo_np = rw.diffract(oe1, wave_np)
    rw._diffraction_integral_conv = orig
    # OpenCL branch of the unmodified rw.diffract with the recorder at the boundary
    rec = Recorder(lambda: conv['res'])
    rw.waveCL = rec
    try:
        wave_cl = scr.prepare_wave(slit, xm, zm)
        oe2 = copy.deepcopy(oeLocal)
        glo_cl = rw.diffract(oe2, wave_cl)
    finally:
        rw.waveCL = None
    assert len(rec.calls) == 1 and rec.calls[0]['kernelName'] == 'integrate_kirchhoff'
    for f in ('Es', 'Ep', 'Jss', 'Jpp', 'a', 'b', 'c'):
        assert np.array_equal(getattr(wave_cl, f), getattr(wave_np, f)), f  # both branches end identically
        assert np.array_equal(getattr(glo_cl, f), getattr(glo_np, f)), f
    pack_call('kir_', rec.calls[0], out)
    for name, v in zip(('Es', 'Ep', 'aE', 'bE', 'cE'), conv['res']):
        out['kir_numpy_' + name] = v
    for f in ('Es', 'Ep', 'Jss', 'Jpp', 'a', 'b', 'c'):
        out['kir_wave_' + f] = getattr(wave_np, f)
    print('diffract: recorded', rec.calls[0]['dimension'], 'receiving points x', int(rec.calls[0]['scalarArgs'][0]),
          'samples')


def record_undulator(out):
    bl = raycing.BeamLine()
    kwargs = dict(eE=3., eI=0.5, period=30., n=40, K=1.45, eSigmaX=48.65, eSigmaZ=6.197, eEpsilonX=0.263,
                  eEpsilonZ=0.008, xPrimeMax=0.1, zPrimeMax=0.1, xPrimeMaxAutoReduce=False,
                  zPrimeMaxAutoReduce=False, targetOpenCL=None, gNodes=32, gIntervals=2)
    E1 = 1388.8614066632863
    src = rs.Undulator(bl, eMin=E1*0.995, eMax=E1*1.005, **kwargs)
    src.reset()
    rng = np.random.RandomState(77)
    n = 1200
    w = rng.uniform(src.E_min, src.E_max, n)
    th = rng.uniform(src.Theta_min, src.Theta_max, n)
    ps = rng.uniform(src.Psi_min, src.Psi_max, n)
    spsum = {}
    orig = rs.Undulator._sp_sum

    def spy(self, *a, **kw):
        res = orig(self, *a, **kw)
        spsum['res'] = [np.array(v) for v in res]
        return res
    rs.Undulator._sp_sum = spy
    I_np, Es_np, Ep_np = src.build_I_map(w, th, ps)  # cl_ctx is None: numpy branch
    rs.Undulator._sp_sum = orig
    rec = Recorder(lambda: spsum['res'])
    src.ucl = rec
    src.cl_ctx = rec.cl_ctx
    src.cl_precisionF = rec.cl_precisionF
    src.cl_precisionC = rec.cl_precisionC
    I_cl, Es_cl, Ep_cl = src.build_I_map(w, th, ps)  # OpenCL branch of the unmodified reference
    assert len(rec.calls) == 1 and rec.calls[0]['kernelName'] == 'undulator'
    # the two branches of the reference agree to rounding when the kernel returns _sp_sum's numbers (ab * Is vs Is * ab)
    assert np.abs(I_cl - I_np).max() <= 1e-14 * np.abs(I_np).max()
    assert np.abs(Es_cl - Es_np).max() <= 1e-14 * np.abs(Es_np).max()
    pack_call('und_', rec.calls[0], out)
    out['und_numpy_Is'], out['und_numpy_Ip'] = spsum['res']
    out['und_I'], out['und_Es'], out['und_Ep'] = I_np, Es_np, Ep_np
    print('undulator: recorded', rec.calls[0]['dimension'], 'samples x', int(rec.calls[0]['scalarArgs'][3]), 'nodes')


if __name__ == '__main__':
    out = {}
    record_diffract(out)
    record_undulator(out)
    path = os.path.join(ROOT, 'tests', 'golden', 'dropin_calls.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes')
