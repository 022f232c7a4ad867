# -*- coding: utf-8 -*-
"""Device-resident `Beam`: the SoA ray container of xrt (xrt/backends/raycing/sources/beams.py:96-186) with
every array living in HBM as a torch CUDA tensor.

Same attribute names (x y z a b c path E Jss Jpp Jsp state [Es Ep] [theta s phi r] ...), same constructor
keywords (`nrays`, `copyFrom`, `withAmplitudes`), same bookkeeping scalars (`accepted`, `seeded`, ...), so
that `run_process()` scripts and the plotting layer read it the same way; `to_numpy()` /
`Beam.from_numpy()` cross the host boundary explicitly (100 B/ray, 132 B/ray with amplitudes).
"""
import ctypes

import numpy as np
import torch

F64 = torch.float64
C128 = torch.complex128
I32 = torch.int32

REAL_FIELDS = ('x', 'y', 'z', 'a', 'b', 'c', 'path', 'E', 'Jss', 'Jpp')
COMPLEX_FIELDS = ('Jsp', 'Es', 'Ep')
OPTIONAL_REAL = ('theta', 's', 'phi', 'r')
SCALARS = ('accepted', 'acceptedE', 'seeded', 'seededI', 'sourceWeight', 'parentId', 'sourceSIGMAx',
           'sourceSIGMAz', 'filamentDX', 'filamentDZ', 'filamentDtheta', 'filamentDpsi', 'filamentDgamma')
# Beam(copyFrom=...) copies exactly the reference's listOfAttrs (sources/beams.py:96-105)
COPIED = ('x', 'y', 'z', 'sourceSIGMAx', 'sourceSIGMAz', 'filamentDX', 'filamentDZ', 'filamentDtheta', 'filamentDpsi',
          'filamentDgamma', 'state', 'a', 'b', 'c', 'path', 'E', 'Jss', 'Jpp', 'Jsp', 'elevationD', 'elevationX',
          'elevationY', 'elevationZ', 's', 'phi', 'r', 'theta', 'order', 'accepted', 'acceptedE', 'seeded', 'seededI',
          'Es', 'Ep', 'nRefl', 'sourceWeight')


class xrt_beam(ctypes.Structure):
    """struct xrt_beam of include/xrt_b200.h"""
    _fields_ = [(n, ctypes.c_void_p) for n in
                ('x', 'y', 'z', 'a', 'b', 'c', 'path', 'E', 'Jss', 'Jpp', 'Jsp', 'Es', 'Ep', 'state',
                 'theta', 's', 'phi', 'r')]


def _device(device=None):
    if device is None:
        return torch.device('cuda', torch.cuda.current_device())
    return torch.device(device)


class _HostTransfer(object):
    """Handle of Beam.to_host_begin: the compacted device arrays, the pinned count and the event that says both are
    final."""
    _copy_streams = {}

    def __init__(self, names, dst, cnt_host, ready, out, scratch):
        self.names, self.dst, self.cnt_host, self.ready, self.out, self._scratch = names, dst, cnt_host, ready, out, scratch

    def finish(self):
        """Waits for the selected count, copies the fields on the copy stream and returns {field: numpy array}."""
        self.ready.synchronize()
        cnt = int(self.cnt_host[0])
        dev = self.dst[0].device if self.dst else torch.device('cuda', torch.cuda.current_device())
        cs = _HostTransfer._copy_streams.get(str(dev))
        if cs is None:
            cs = _HostTransfer._copy_streams[str(dev)] = torch.cuda.Stream(device=dev)
        res = {}
        with torch.cuda.stream(cs):
            for k, d in zip(self.names, self.dst):
                if self.out is not None and k in self.out:
                    self.out[k][:cnt].copy_(d[:cnt], non_blocking=True)
                    res[k] = self.out[k][:cnt]
                else:
                    h = torch.empty(cnt, dtype=d.dtype, pin_memory=True)
                    h.copy_(d[:cnt], non_blocking=True)
                    res[k] = h
                d.record_stream(cs)
        cs.synchronize()
        self.dst = self._scratch = None
        return {k: v.numpy() for k, v in res.items()}


class Beam(object):
    def __init__(self, nrays=0, copyFrom=None, withAmplitudes=False, device=None, uninitialized=False):
        if copyFrom is not None:  # only the reference's listOfAttrs travel (beams.py:96-117): no area, toOE, ...
            for k in COPIED:
                if hasattr(copyFrom, k):
                    v = getattr(copyFrom, k)
                    setattr(self, k, v.clone() if isinstance(v, torch.Tensor) else v)
            self.parentId = getattr(copyFrom, 'parentId', None)
            return
        dev = _device(device)
        n = int(nrays)
        new = torch.empty if uninitialized else torch.zeros
        for f in REAL_FIELDS:
            setattr(self, f, new(n, dtype=F64, device=dev))
        self.Jsp = new(n, dtype=C128, device=dev)
        self.state = new(n, dtype=I32, device=dev)
        if withAmplitudes:
            self.Es = new(n, dtype=C128, device=dev)
            self.Ep = new(n, dtype=C128, device=dev)
        if not uninitialized:
            # defaults of the reference constructor (beams.py:165-182)
            self.b.fill_(1.)
            self.E.fill_(9000.)
            self.Jss.fill_(1.)
        self.parentId = None

    # ------------------------------------------------------------------ structure
    def __len__(self):
        return self.x.numel()

    @property
    def nrays(self):
        return self.x.numel()

    @property
    def device(self):
        return self.x.device

    def has_amplitudes(self):
        return hasattr(self, 'Es')

    def like(self, withLocal=False):
        """Uninitialised beam of the same shape (output buffer of a kernel)."""
        b = Beam(len(self), withAmplitudes=self.has_amplitudes(), device=self.device, uninitialized=True)
        for k in SCALARS:
            if hasattr(self, k):
                setattr(b, k, getattr(self, k))
        return b

    def tensors(self):
        return {k: v for k, v in self.__dict__.items() if isinstance(v, torch.Tensor)}

    def bytes_per_ray(self):
        return sum(v.element_size() for v in self.tensors().values())

    def c_struct(self):
        s = xrt_beam()
        for name, _ in xrt_beam._fields_:
            t = getattr(self, name, None)
            if isinstance(t, torch.Tensor):
                if not (t.is_cuda and t.is_contiguous() and t.numel() == len(self)):
                    raise ValueError('Beam.{0} must be a contiguous CUDA tensor of the beam length'.format(name))
                setattr(s, name, t.data_ptr())
            else:
                setattr(s, name, None)
        return s

    # ------------------------------------------------------------------ host boundary
    @classmethod
    def from_numpy(cls, arrays, device=None, pinned=False):
        """arrays: mapping name -> ndarray (or an object with those attributes, e.g. a reference Beam)."""
        if not isinstance(arrays, dict):
            arrays = {k: v for k, v in arrays.__dict__.items()}
        dev = _device(device)
        b = cls.__new__(cls)
        for k, v in arrays.items():
            if isinstance(v, np.ndarray) and v.ndim == 1:
                if k == 'state':
                    v = np.ascontiguousarray(v, dtype=np.int32)
                elif k in COMPLEX_FIELDS:
                    v = np.ascontiguousarray(v, dtype=np.complex128)
                else:
                    v = np.ascontiguousarray(v, dtype=np.float64)
                t = torch.from_numpy(v)
                setattr(b, k, t.to(dev, non_blocking=pinned))
            elif k in SCALARS or k in ('area', 'createdByDiffract', 'dS', 'areaNormal'):
                setattr(b, k, v)
        if not hasattr(b, 'parentId'):
            b.parentId = None
        return b

    def to_numpy(self):
        out = {}
        for k, v in self.__dict__.items():
            out[k] = v.cpu().numpy() if isinstance(v, torch.Tensor) else v
        return out

    _STATE_BITS = {1: 1, 2: 2, 3: 4, 'lost': 8, 'other': 16, 0: 32}

    def to_host(self, fields=('x', 'z', 'E', 'Jss', 'Jpp'), good_only=True, rayFlag=(1,), out=None, stream=None):
        """The fields a consumer of a repeat reads, of the rays it reads them for, as numpy arrays -- and only those
        cross PCIe.  raycing.get_output (raycing/__init__.py:170-300) takes a few attributes of the rays whose state
        is in the plot's rayFlag; moving the whole SoA (132 B x every ray) to the host for that is ~5x the bytes.

        *good_only* / *rayFlag*: which states are kept (1 good, 2 out, 3 over, 4 lost = negative states), in ray
        order; good_only=False keeps every ray (no compaction, plain copies of the chosen fields).
        *out*: optional dict of pinned host tensors to copy into (a reusable result buffer, each at least as long as
        the selection); returned views are cut to the selected length.
        Device side: xrt_compact_rays_f64 (count, scan, scatter), stream-ordered; one 8-byte read tells the host
        the selected count, then one copy per field."""
        import ctypes
        from . import _lib
        lib = _lib.load()
        n = len(self)
        dev = self.device
        names = list(fields)
        tens = [getattr(self, k) for k in names]
        if not good_only:
            res = {}
            for k, t in zip(names, tens):
                if out is not None and k in out:
                    out[k][:n].copy_(t, non_blocking=True)
                    res[k] = out[k][:n]
                else:
                    res[k] = t.cpu()
            torch.cuda.current_stream().synchronize()
            return {k: v.numpy() for k, v in res.items()}
        mask = 0
        for f in rayFlag:
            mask |= {1: 1, 2: 2, 3: 4, 4: 8}[int(f)]
        if n == 0:
            return {k: t.cpu().numpy() for k, t in zip(names, tens)}
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        key = '_compact_scratch'
        scratch = self.__dict__.get(key)
        need = lib.xrt_compact_scratch_words(n)
        if scratch is None or scratch.numel() < need:
            scratch = torch.empty(need, dtype=torch.int64, device=dev)
        dst = [torch.empty_like(t) for t in tens]
        res = {}
        for c0 in range(0, len(names), 8):
            grp = list(range(c0, min(c0 + 8, len(names))))
            m = len(grp)
            src_p = (ctypes.c_void_p * m)(*[tens[i].data_ptr() for i in grp])
            dst_p = (ctypes.c_void_p * m)(*[dst[i].data_ptr() for i in grp])
            words = (ctypes.c_int * m)(*[2 if tens[i].is_complex() else 1 for i in grp])
            for i in grp:
                if tens[i].element_size() not in (8, 16) or not tens[i].is_contiguous():
                    raise ValueError('to_host: {0} is not a contiguous float64 / complex128 array'.format(names[i]))
            _lib.check(lib.xrt_compact_rays_f64(n, self.state.data_ptr(), mask, m, ctypes.addressof(src_p),
                                                ctypes.addressof(dst_p), ctypes.addressof(words),
                                                scratch.data_ptr(), st), 'xrt_compact_rays_f64')
        cnt = int(scratch[need - 1].item())  # the one host synchronisation: how many rays were selected
        for k, d in zip(names, dst):
            if out is not None and k in out:
                out[k][:cnt].copy_(d[:cnt], non_blocking=True)
                res[k] = out[k][:cnt]
            else:
                res[k] = d[:cnt].cpu()
        torch.cuda.current_stream().synchronize()
        return {k: v.numpy() for k, v in res.items()}

    def to_host_begin(self, fields=('x', 'z', 'E', 'Jss', 'Jpp'), rayFlag=(1,), out=None):
        """First half of an ASYNCHRONOUS to_host(good_only=True): launches the compaction on the current stream and an
        8-byte copy of the selected count into pinned memory, then returns a handle at once.  `handle.finish()` -- called
        after the NEXT step's kernels have been launched -- waits for the count, moves the compacted fields on a copy
        stream of its own (PCIe runs while the GPU computes the next step) and returns the numpy views.  *out*: dict of
        pinned host tensors, one per field (alternate between two sets when handles overlap)."""
        import ctypes
        from . import _lib
        lib = _lib.load()
        n = len(self)
        names = list(fields)
        tens = [getattr(self, k) for k in names]
        mask = 0
        for f in rayFlag:
            mask |= {1: 1, 2: 2, 3: 4, 4: 8}[int(f)]
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        need = lib.xrt_compact_scratch_words(max(n, 1))
        scratch = torch.empty(need, dtype=torch.int64, device=self.device)
        dst = [torch.empty_like(t) for t in tens]
        for c0 in range(0, len(names), 8):
            grp = list(range(c0, min(c0 + 8, len(names))))
            m = len(grp)
            for i in grp:
                if tens[i].element_size() not in (8, 16) or not tens[i].is_contiguous():
                    raise ValueError('to_host: {0} is not a contiguous float64 / complex128 array'.format(names[i]))
            src_p = (ctypes.c_void_p * m)(*[tens[i].data_ptr() for i in grp])
            dst_p = (ctypes.c_void_p * m)(*[dst[i].data_ptr() for i in grp])
            words = (ctypes.c_int * m)(*[2 if tens[i].is_complex() else 1 for i in grp])
            if n:
                _lib.check(lib.xrt_compact_rays_f64(n, self.state.data_ptr(), mask, m, ctypes.addressof(src_p),
                                                    ctypes.addressof(dst_p), ctypes.addressof(words),
                                                    scratch.data_ptr(), st), 'xrt_compact_rays_f64')
        cnt_host = torch.zeros(1, dtype=torch.int64, pin_memory=True)
        if n:
            cnt_host.copy_(scratch[need - 1:need], non_blocking=True)
        ready = torch.cuda.Event()
        ready.record()
        return _HostTransfer(names, dst, cnt_host, ready, out, scratch)

    def __array__(self):  # pragma: no cover - convenience for interactive use
        raise TypeError('Beam is a container; use beam.to_numpy()')

    def filter_by_index(self, index):
        for k, v in list(self.__dict__.items()):
            if isinstance(v, torch.Tensor):
                setattr(self, k, v[index].contiguous())

    def concatenate(self, other):
        for k, v in list(self.__dict__.items()):
            if isinstance(v, torch.Tensor):
                setattr(self, k, torch.cat((v, getattr(other, k))))
