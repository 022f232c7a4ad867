# -*- coding: utf-8 -*-
"""`XRT_CUDA` -- drop-in for xrt's accelerator shim `XRT_CL`.

Mirrors the object surface the reference's call sites use (xrt/backends/raycing/myopencl.py:86-97 ctor,
:155-412 `set_cl`, :414-583 `run_parallel`): same constructor arguments, same `run_parallel` keyword
signature and argument convention (scalars, sliced RO, non-sliced RO, sliced RW, non-sliced RW; RW arrays are
written back into the caller's arrays *and* returned), same attributes (`cl_precisionF/C`, `cl_ctx`,
`lastTargetOpenCL`, `cl_is_blocking`).  Kernels are looked up by the reference's kernel names:

    'integrate_kirchhoff'  cl/diffract.cl:80     called from waves.py:892
    'undulator'            cl/undulator.cl:54    called from sources/synchr.py:2174
    'get_amplitude'        cl/materials.cl:715   called from tests/raycing/test_materials.py:624
    'get_amplitude_graded_multilayer[_tran]'  cl/materials.cl:568-840  called from materials/multilayer.py:471-503
    'undulator_taper', 'undulator_nf', 'custom_field'  cl/undulator.cl:123-298  called from sources/synchr.py:2174

Differences, on purpose: fp64 only (`precisionOpenCL` 'auto'/'float64'); there is no silent `_disable_cl`
fallback to numpy -- a missing library or GPU raises; multi-device slicing (`run_parallel_max`,
myopencl.py:455-583) is replaced by one process per GPU (see xrt_b200/parallel.py).
"""
import numpy as np

from . import _lib, kernels

isOpenCL = False  # there is no OpenCL here; kept so `mcl.isOpenCL or mcl.isZMQ` style tests read the same
isZMQ = False
isCUDA = True
ALL_CL_DEVICES = {'None': None, 'auto': 'auto', 'GPU': 'GPU', 'CUDA': 'CUDA'}

_KIRCHHOFF_CONVENTIONS = {'numpy': 0, 'opencl': 1}


class XRT_CUDA(object):
    def __init__(self, filename=None, targetOpenCL='auto', precisionOpenCL='auto', kernelsource=None):
        self.kernelsource = kernelsource  # accepted for signature parity; CUDA kernels are precompiled
        self.cl_filename = filename
        self.lastTargetOpenCL = None
        self.lastPrecisionOpenCL = None
        self.lastOpenCLError = None
        self.cl_is_blocking = True
        self.cl_queue = []
        self.cl_ctx = []
        self.cl_program = []
        self.cl_mf = None
        self.cl_precisionF = None
        self.cl_precisionC = None
        self.useZMQ = False
        #: sign/scale convention of 'integrate_kirchhoff': 'numpy' follows waves._diffraction_integral_conv
        #: (waves.py:844-850), 'opencl' follows cl/diffract.cl:137-149.  They differ by constant phase factors
        #: that cancel in waves.diffract (last paragraph).
        self.kirchhoffConvention = 'numpy'
        self.kirchhoffRelaxed = False
        self.set_cl(targetOpenCL, precisionOpenCL)

    # ------------------------------------------------------------------ configuration
    def set_cl(self, targetOpenCL='auto', precisionOpenCL='auto', strict=True):
        if (targetOpenCL == self.lastTargetOpenCL) and (precisionOpenCL == self.lastPrecisionOpenCL):
            return
        if targetOpenCL is None:
            self.lastTargetOpenCL = None
            self.cl_ctx = []
            return
        device = None
        if isinstance(targetOpenCL, str):
            t = targetOpenCL.strip()
            if t.upper() in ('AUTO', 'GPU', 'CUDA', 'ALL'):
                device = None
            elif t.upper().startswith('CUDA:'):
                device = int(t.split(':', 1)[1].split(',')[0])
            else:
                raise ValueError("XRT_CUDA: unsupported targetOpenCL {0!r}; use 'auto', 'GPU', 'CUDA' or "
                                 "'CUDA:<index>' (there is no CPU or OpenCL path)".format(targetOpenCL))
        elif isinstance(targetOpenCL, (tuple, list)) and len(targetOpenCL) == 2 and \
                all(isinstance(v, (int, np.integer)) for v in targetOpenCL):
            device = int(targetOpenCL[1])  # (iPlatform, iDevice) of the reference -> CUDA device index
        elif isinstance(targetOpenCL, (int, np.integer)):
            device = int(targetOpenCL)
        else:
            raise ValueError("XRT_CUDA: unsupported targetOpenCL {0!r}".format(targetOpenCL))
        if precisionOpenCL not in ('auto', 'float64'):
            raise ValueError("XRT_CUDA computes in float64 only (got precisionOpenCL={0!r})".format(
                precisionOpenCL))
        self.device = _lib.init(device)  # raises without a GPU: no CPU fallback
        self.lastTargetOpenCL = targetOpenCL
        self.lastPrecisionOpenCL = precisionOpenCL
        self.cl_precisionF = np.float64
        self.cl_precisionC = np.complex128
        self.cl_ctx = ['cuda:{0}'.format(self.device)]  # truthiness is what callers test (synchr.py:2043)
        self.cl_queue = [None]
        self.cl_program = [None]

    # ------------------------------------------------------------------ dispatch
    def run_parallel(self, kernelName='', scalarArgs=None, slicedROArgs=None, nonSlicedROArgs=None,
                     slicedRWArgs=None, nonSlicedRWArgs=None, dimension=0, complexity=0, signal=None):
        if self.lastTargetOpenCL is None:
            raise _lib.XrtB200Error('XRT_CUDA is disabled (targetOpenCL=None)')
        try:
            fn = getattr(self, '_k_' + kernelName)
        except AttributeError:
            raise NotImplementedError(
                "XRT_CUDA has no kernel named {0!r}; hot-path kernels: {1}".format(
                    kernelName, sorted(n[3:] for n in dir(self) if n.startswith('_k_'))))
        return fn(scalarArgs or [], slicedROArgs or [], nonSlicedROArgs or [], slicedRWArgs or [],
                  nonSlicedRWArgs or [], int(dimension))

    #: a xrt_b200.parallel.Comm: run_parallel_max then shards over the ranks (one process per GPU)
    comm = None

    def run_parallel_max(self, kernelName='', scalarArgs=None, slicedROArgs=None, nonSlicedROArgs=None,
                         slicedRWArgs=None, nonSlicedRWArgs=None, dimension=0, complexity=0, signal=None):
        """XRT_CL.run_parallel_max (myopencl.py:455-583) over ranks instead of over the OpenCL devices of one
        process: every rank calls it with the SAME whole host arrays (SPMD); the sliced arguments are cut into
        contiguous chunks of `dimension` (x secondDim, :501-503) -- this rank computes its chunk --, the non-sliced
        ones are replicated (:508-512), and the sliced RW results of all ranks are gathered back into the caller's
        whole arrays (:552-569), which are also returned.  Without a communicator it is run_parallel."""
        comm = self.comm
        if comm is None or comm.world == 1:
            return self.run_parallel(kernelName, scalarArgs, slicedROArgs, nonSlicedROArgs, slicedRWArgs,
                                     nonSlicedRWArgs, dimension, complexity, signal)
        if nonSlicedRWArgs:
            raise NotImplementedError('run_parallel_max: non-sliced RW arguments cannot be sharded')
        dimension = int(dimension)
        start, stop = comm.shard(dimension)

        def cut(a):
            a = np.asarray(a)
            second = len(a) // dimension if dimension else 1  # secondDim (myopencl.py:501-503)
            return a[start * second:stop * second], second
        ro = [cut(a)[0] for a in (slicedROArgs or [])]
        rw_full = list(slicedRWArgs or [])
        rw = []
        for a in rw_full:
            part, _ = cut(a)
            rw.append(np.ascontiguousarray(part).copy())
        self.run_parallel(kernelName, scalarArgs, ro, nonSlicedROArgs, rw, None, stop - start, complexity, signal)
        for full, part in zip(rw_full, rw):
            second = len(full) // dimension if dimension else 1
            g = comm.allgather_host(part.reshape(stop - start, second) if second > 1 else part,
                                    device=getattr(self, '_gather_device', None))
            full[...] = g.reshape(full.shape)
        return tuple(rw_full)

    # ------------------------------------------------------------------ kernels
    def _k_integrate_kirchhoff(self, scalarArgs, slicedRO, nonSlicedRO, slicedRW, nonSlicedRW, dimension):
        """Argument layout of waves._diffraction_integral_CL (waves.py:864-894)."""
        lib = _lib.load()
        N = int(scalarArgs[0])
        xM, yM, zM = (_lib.as_f64(a) for a in slicedRO)
        M = dimension if dimension else len(xM)
        nl, Es, Ep, k, pos, nrm = nonSlicedRO
        nl = _lib.as_f64(nl)
        Es = _lib.as_c128(Es)
        Ep = _lib.as_c128(Ep)
        k = _lib.as_f64(k)
        pos, pstride = _as_vec_buffer(pos, N)
        nrm, nstride = _as_vec_buffer(nrm, N)
        if not (len(nl) == len(Es) == len(Ep) == len(k) == N):
            raise ValueError('integrate_kirchhoff: non-sliced inputs must all have length fullnrays')
        outs = []
        for a in slicedRW:
            if not (isinstance(a, np.ndarray) and a.dtype == np.complex128 and a.flags['C_CONTIGUOUS'] and
                    len(a) == M):
                raise ValueError('integrate_kirchhoff: slicedRWArgs must be 5 contiguous complex128[dimension]')
            outs.append(a)
        flags = (1 if self.kirchhoffRelaxed else 0) | (_KIRCHHOFF_CONVENTIONS[self.kirchhoffConvention] << 1)
        if N > 0 and M > 0:
            # sin / cos through the shared-memory table (same accuracy, faster) when every phase k r stays below
            # kernels.PHASE_LIMIT_TABLE_SINCOS:
            # bound it with the bounding boxes of the samples and of the receiving points
            pv = pos[:3, :] if pos.shape[1] == N and pos.shape[0] in (3, 4) and pos.flags['F_CONTIGUOUS'] else pos[:, :3].T
            span = np.maximum(np.abs(np.array([xM.max(), yM.max(), zM.max()]) - pv.min(axis=1)),
                              np.abs(pv.max(axis=1) - np.array([xM.min(), yM.min(), zM.min()])))
            if np.abs(k).max() * np.linalg.norm(span) < kernels.PHASE_LIMIT_TABLE_SINCOS:
                flags |= 1 << 11
        nview = nrm[:3, :] if nrm.shape[1] == N and nrm.shape[0] in (3, 4) and nrm.flags['F_CONTIGUOUS'] else nrm[:, :3].T
        if N > 0 and np.all(nview == nview[:, :1]):
            flags |= 1 << 9  # all normals equal (aperture / flat OE): the kernel hoists n.p out of the pair loop
        rc = lib.xrt_kirchhoff_f64_host(
            M, N, _lib.hptr(xM), _lib.hptr(yM), _lib.hptr(zM), _lib.hptr(nl), _lib.hptr(Es), _lib.hptr(Ep),
            _lib.hptr(k), _lib.hptr(pos), pstride, _lib.hptr(nrm), nstride,
            *[_lib.hptr(a) for a in outs], flags)
        _lib.check(rc, 'xrt_kirchhoff_f64_host')
        return tuple(outs)


    def _k_undulator(self, scalarArgs, slicedRO, nonSlicedRO, slicedRW, nonSlicedRW, dimension):
        """Argument layout of Undulator._build_I_map_CL (sources/synchr.py:2134-2155):
        scalars [alpha, Kx, Ky, jend]; sliced RO [gamma, wu, w, ww1, ddtheta, ddpsi];
        non-sliced RO [tg, ag, sintg, costg, sintgph, costgph]; sliced RW [Is, Ip]."""
        lib = _lib.load()
        alpha, Kx, Ky, jend = scalarArgs[:4]
        samp = [_lib.as_f64(a) for a in slicedRO]
        n = dimension if dimension else len(samp[0])
        tabs = [_lib.as_f64(a) for a in nonSlicedRO]
        if any(len(a) != n for a in samp) or len(samp) != 6:
            raise ValueError('undulator: expected 6 sliced inputs of length dimension')
        if any(len(a) != int(jend) for a in tabs) or len(tabs) != 6:
            raise ValueError('undulator: expected 6 node tables of length jend')
        Is, Ip = slicedRW
        for a in (Is, Ip):
            if not (isinstance(a, np.ndarray) and a.dtype == np.complex128 and a.flags['C_CONTIGUOUS'] and
                    len(a) == n):
                raise ValueError('undulator: slicedRWArgs must be 2 contiguous complex128[dimension]')
        rc = lib.xrt_undulator_f64_host(
            n, int(jend), float(alpha), float(Kx), float(Ky), *[_lib.hptr(a) for a in samp],
            *[_lib.hptr(a) for a in tabs], _lib.hptr(Is), _lib.hptr(Ip), 0)
        _lib.check(rc, 'xrt_undulator_f64_host')
        return Is, Ip


    def _undulator_ext(self, variant, name, scalarArgs, slicedRO, nonSlicedRO, slicedRW, dimension):
        """Argument layout of Undulator._build_I_map_CL for the tapered / near-field kernels
        (sources/synchr.py:2134-2176): scalars [taper | R0*2pi/L0, Kx, Ky, jend, Np]; arrays as `undulator`."""
        lib = _lib.load()
        if len(scalarArgs) < 5:
            raise ValueError(name + ': expected 5 scalars [taper or R0*2pi/L0, Kx, Ky, jend, Np]')
        scalar0, Kx, Ky, jend, nmax = scalarArgs[:5]
        samp = [_lib.as_f64(a) for a in slicedRO]
        n = dimension if dimension else len(samp[0])
        tabs = [_lib.as_f64(a) for a in nonSlicedRO]
        if any(len(a) != n for a in samp) or len(samp) != 6:
            raise ValueError(name + ': expected 6 sliced inputs of length dimension')
        if any(len(a) != int(jend) for a in tabs) or len(tabs) != 6:
            raise ValueError(name + ': expected 6 node tables of length jend')
        Is, Ip = slicedRW
        for a in (Is, Ip):
            if not (isinstance(a, np.ndarray) and a.dtype == np.complex128 and a.flags['C_CONTIGUOUS'] and
                    len(a) == n):
                raise ValueError(name + ': slicedRWArgs must be 2 contiguous complex128[dimension]')
        rc = lib.xrt_undulator_ext_f64_host(
            variant, n, int(jend), int(nmax), float(scalar0), float(Kx), float(Ky), *[_lib.hptr(a) for a in samp],
            *[_lib.hptr(a) for a in tabs], _lib.hptr(Is), _lib.hptr(Ip))
        _lib.check(rc, 'xrt_undulator_ext_f64_host')
        return Is, Ip

    def _k_undulator_taper(self, scalarArgs, slicedRO, nonSlicedRO, slicedRW, nonSlicedRW, dimension):
        """kernel `undulator_taper` (cl/undulator.cl:123-202), numpy-path numerics (synchr.py:1958-1971)."""
        return self._undulator_ext(1, 'undulator_taper', scalarArgs, slicedRO, nonSlicedRO, slicedRW, dimension)

    def _k_undulator_nf(self, scalarArgs, slicedRO, nonSlicedRO, slicedRW, nonSlicedRW, dimension):
        """kernel `undulator_nf` (cl/undulator.cl:204-298), numpy-path numerics (synchr.py:1972-1996)."""
        return self._undulator_ext(2, 'undulator_nf', scalarArgs, slicedRO, nonSlicedRO, slicedRW, dimension)

    def _k_custom_field(self, scalarArgs, slicedRO, nonSlicedRO, slicedRW, nonSlicedRW, dimension):
        """Argument layout of SourceFromField._build_I_map_custom_field_CL (sources/synchr.py:1227-1259): scalars
        [jend, betam, R0]; sliced RO [gamma, w, ddtheta, ddpsi]; non-sliced RO [tg, ag, Bx, By, Bz, betax, betay,
        trajx, trajy, trajz]; sliced RW [Is, Ip]."""
        lib = _lib.load()
        jend, betam, R0 = scalarArgs[:3]
        samp = [_lib.as_f64(a) for a in slicedRO]
        n = dimension if dimension else len(samp[0])
        tabs = [_lib.as_f64(a) for a in nonSlicedRO]
        if len(samp) != 4 or any(len(a) != n for a in samp):
            raise ValueError('custom_field: expected 4 sliced inputs of length dimension')
        if len(tabs) != 10 or any(len(a) != int(jend) for a in tabs):
            raise ValueError('custom_field: expected 10 node tables of length jend')
        Is, Ip = slicedRW
        for a in (Is, Ip):
            if not (isinstance(a, np.ndarray) and a.dtype == np.complex128 and a.flags['C_CONTIGUOUS'] and
                    len(a) == n):
                raise ValueError('custom_field: slicedRWArgs must be 2 contiguous complex128[dimension]')
        rc = lib.xrt_custom_field_f64_host(n, int(jend), float(betam), float(R0) if R0 is not None else 0.,
                                           *[_lib.hptr(a) for a in samp], *[_lib.hptr(a) for a in tabs],
                                           _lib.hptr(Is), _lib.hptr(Ip))
        _lib.check(rc, 'xrt_custom_field_f64_host')
        return Is, Ip

    def _k_get_amplitude(self, scalarArgs, slicedRO, nonSlicedRO, slicedRW, nonSlicedRW, dimension):
        """Argument list of kernel `get_amplitude` (cl/materials.cl:715-748): scalars [hkl (int4), factDW, thickness
        (mm), geom_b, maxEl]; sliced RO [BeamInDotNormal, BeamOutDotNormal, BeamInDotHNormal, Energy, IPDistance (d,
        Angstrom), ChiToF]; non-sliced RO [elements (maxEl x 8: x, y, z, -, occupancy, Z, -, table length), f0cfs
        (maxEl x 11), E_vector, f1_vector, f2_vector (maxEl x 300)]; sliced RW [refl_s, refl_p].
        geom_b as the kernel decodes it (:734-736): Bragg when > 1, transmitted when odd; thickness 0 or > 0.5 mm with
        'Bragg reflected' is the thick-crystal branch (:268).  Numerics are those of the numpy path
        (Crystal.get_amplitude, crystal.py:492-645 -- SURVEY 8a: numpy is the oracle).  The cell may hold any number of
        atoms of ONE element (fcc, diamond, ... : their geometric structure factor is taken on the host); cells with
        different elements are outside the hot path."""
        import ctypes
        from .materials import xrt_material_params
        lib = _lib.load()
        hkl, factDW, thickness, geom_b, maxEl = scalarArgs[:5]
        hkl = np.asarray(hkl, dtype=np.float64).ravel()[:3]
        maxEl = int(maxEl)
        geom_b = int(geom_b)
        thickness = float(thickness)
        bIn, bOut, bH, E, dArr, chiArr = (_lib.as_f64(a) for a in slicedRO)
        n = dimension if dimension else len(E)
        if any(len(a) != n for a in (bIn, bOut, bH, E, dArr, chiArr)):
            raise ValueError('get_amplitude: expected 6 sliced inputs of length dimension')
        el, f0cfs, Ev, f1v, f2v = (np.asarray(a, dtype=np.float64).ravel() for a in nonSlicedRO)
        if maxEl < 1 or len(el) < 8 * maxEl or len(f0cfs) < 11 * maxEl or min(len(Ev), len(f1v), len(f2v)) < 300 * maxEl:
            raise ValueError('get_amplitude: elements / f0cfs / E, f1, f2 vectors are shorter than maxEl needs')
        el = el[:8 * maxEl].reshape(maxEl, 8)
        nt = int(round(el[0, 7]))
        same = all(el[i, 5] == el[0, 5] and int(round(el[i, 7])) == nt and
                   np.array_equal(f0cfs[11*i:11*i + 11], f0cfs[:11]) and
                   np.array_equal(Ev[300*i:300*i + nt], Ev[:nt]) and np.array_equal(f1v[300*i:300*i + nt], f1v[:nt])
                   and np.array_equal(f2v[300*i:300*i + nt], f2v[:nt]) for i in range(maxEl))
        if not same:
            raise NotImplementedError('get_amplitude: cells with different elements are outside the hot path '
                                      '(one-element structures: fcc, diamond, ...)')
        rs, rp = slicedRW
        for a in (rs, rp):
            if not (isinstance(a, np.ndarray) and a.dtype == np.complex128 and a.flags['C_CONTIGUOUS'] and
                    len(a) == n):
                raise ValueError('get_amplitude: slicedRWArgs must be 2 contiguous complex128[dimension]')
        occ = el[:, 4]
        S = np.sum(occ * np.exp(2j * np.pi * (el[:, :3] @ hkl)))  # sum o_i exp(2 pi i hkl.r_i) (materials.cl:213-216)
        tables = np.ascontiguousarray(np.concatenate([Ev[:nt], f1v[:nt], f2v[:nt]]))
        bragg, transmitted = geom_b > 1, (geom_b % 2) == 1
        thick = bragg and not transmitted and (thickness == 0 or thickness > 0.5)
        # d and chiToF arrive per ray; the rays of one crystal share them: one launch per distinct pair
        pairs = np.stack([dArr, chiArr], axis=1)
        uniq, inv = np.unique(pairs, axis=0, return_inverse=True)
        inv = inv.ravel()
        for iu, (d, chiToF) in enumerate(uniq):
            sel = slice(None) if len(uniq) == 1 else np.nonzero(inv == iu)[0]
            m = xrt_material_params()
            m.kind = 3
            m.nelem = 1
            m.Z[0] = float(el[0, 5])
            m.quantity[0] = 1.
            m.tab_off[0], m.tab_len[0] = 0, nt
            m.d, m.chiToF, m.factDW = float(d), float(chiToF), float(factDW)
            m.t = 0. if thick else thickness
            m.geom = (0 if bragg else 1) | (2 if transmitted else 0)
            q = 0.5 / float(d)
            m.f0c[0] = float(f0cfs[5] + np.sum(f0cfs[:5] * np.exp(-f0cfs[6:11] * q * q)))  # get_f0 (materials.cl:46-54)
            m.hkl_allowed = 1
            m.f0scale = float(occ.sum()) / 4.
            m.d2f_re, m.d2f_im = float(S.real) / 4., float(S.imag) / 4.
            args = [np.ascontiguousarray(a[sel]) for a in (E, bIn, bOut, bH)]
            o_s = np.empty(len(args[0]), dtype=np.complex128)
            o_p = np.empty(len(args[0]), dtype=np.complex128)
            rc = lib.xrt_get_amplitude_f64_host(len(args[0]), ctypes.byref(m), _lib.hptr(tables), nt,
                                                *[_lib.hptr(a) for a in args], _lib.hptr(o_s), _lib.hptr(o_p))
            if rc == 3:
                raise ValueError('E is out of the f1/f2 data table range')
            _lib.check(rc, 'xrt_get_amplitude_f64_host')
            rs[sel] = o_s
            rp[sel] = o_p
        return rs, rp

    def _k_get_amplitude_graded_multilayer(self, scalarArgs, slicedRO, nonSlicedRO, slicedRW, nonSlicedRW,
                                           dimension):
        """Argument layout of Multilayer.get_amplitude's accelerator call (materials/multilayer.py:471-484):
        scalars [nPairs]; sliced RO [rbs_s, rbs_p, rtb_s, rtb_p, rvt_s, rvt_p, Qt, Qb] complex; non-sliced RO
        [dti, dbi]; sliced RW [ri_s, ri_p]."""
        lib = _lib.load()
        npairs = int(scalarArgs[0])
        ins = [_lib.as_c128(a) for a in slicedRO]
        n = dimension if dimension else len(ins[0])
        if len(ins) != 8 or any(len(a) != n for a in ins):
            raise ValueError('get_amplitude_graded_multilayer: expected 8 sliced complex inputs of length dimension')
        dti, dbi = (_lib.as_f64(a) for a in nonSlicedRO)
        if len(dti) < npairs or len(dbi) < npairs:
            raise ValueError('get_amplitude_graded_multilayer: dti and dbi must hold nPairs thicknesses')
        ri_s, ri_p = slicedRW
        for a in (ri_s, ri_p):
            if not (isinstance(a, np.ndarray) and a.dtype == np.complex128 and a.flags['C_CONTIGUOUS'] and
                    len(a) == n):
                raise ValueError('get_amplitude_graded_multilayer: slicedRWArgs must be 2 contiguous '
                                 'complex128[dimension]')
        rc = lib.xrt_multilayer_f64_host(n, npairs, *[_lib.hptr(a) for a in ins], _lib.hptr(dti), _lib.hptr(dbi),
                                         _lib.hptr(ri_s), _lib.hptr(ri_p))
        _lib.check(rc, 'xrt_multilayer_f64_host')
        return ri_s, ri_p

    def _k_get_amplitude_graded_multilayer_tran(self, scalarArgs, slicedRO, nonSlicedRO, slicedRW, nonSlicedRW,
                                                dimension):
        """Argument layout of Multilayer.get_amplitude's accelerator call for geom 'transmitted'
        (materials/multilayer.py:485-503): scalars [nPairs, substThickness]; sliced RO: 23 complex arrays (rvt, tvt,
        rbs, tbs, rsv, tsv, rbt, tbt, rtb, ttb for s and p, then Qt, Qb, Qs); non-sliced RO [dti, dbi]; sliced RW
        [ti_s, ti_p]."""
        import ctypes
        lib = _lib.load()
        npairs, substThickness = int(scalarArgs[0]), float(scalarArgs[1])
        ins = [_lib.as_c128(a) for a in slicedRO]
        n = dimension if dimension else len(ins[0])
        if len(ins) != 23 or any(len(a) != n for a in ins):
            raise ValueError('get_amplitude_graded_multilayer_tran: expected 23 sliced complex inputs of length '
                             'dimension')
        dti, dbi = (_lib.as_f64(a) for a in nonSlicedRO)
        if len(dti) < npairs or len(dbi) < npairs:
            raise ValueError('get_amplitude_graded_multilayer_tran: dti and dbi must hold nPairs thicknesses')
        ti_s, ti_p = slicedRW
        for a in (ti_s, ti_p):
            if not (isinstance(a, np.ndarray) and a.dtype == np.complex128 and a.flags['C_CONTIGUOUS'] and
                    len(a) == n):
                raise ValueError('get_amplitude_graded_multilayer_tran: slicedRWArgs must be 2 contiguous '
                                 'complex128[dimension]')
        table = (ctypes.c_void_p * 23)(*[_lib.hptr(a) for a in ins])
        rc = lib.xrt_multilayer_tran_f64_host(n, npairs, substThickness, ctypes.addressof(table), _lib.hptr(dti),
                                              _lib.hptr(dbi), _lib.hptr(ti_s), _lib.hptr(ti_p))
        _lib.check(rc, 'xrt_multilayer_tran_f64_host')
        return ti_s, ti_p


def _as_vec_buffer(a, N):
    """The reference passes float3 buffers as np.array([x, y, z, 0], order='F') of shape (4, N), i.e. N
    records of 4 doubles (waves.py:874-881).  Accept that, or a C-ordered (N, 3|4) array."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 2 and a.shape[1] == N and a.shape[0] in (3, 4) and a.flags['F_CONTIGUOUS']:
        return a, a.shape[0]
    if a.ndim == 2 and a.shape[0] == N and a.shape[1] in (3, 4):
        a = np.ascontiguousarray(a)
        return a, a.shape[1]
    if a.ndim == 2 and a.shape[1] == N and a.shape[0] in (3, 4):
        a = np.asfortranarray(a)
        return a, a.shape[0]
    raise ValueError('expected a (4, N) Fortran-ordered or (N, 3|4) C-ordered float64 array')
