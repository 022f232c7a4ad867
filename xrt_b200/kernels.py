# -*- coding: utf-8 -*-
"""Device-resident entry points: torch CUDA tensors in, torch CUDA tensors out; every call goes through the
C ABI (include/xrt_b200.h) with raw `data_ptr()` addresses on torch's current stream.  torch is used for
device memory, streams and torch.distributed only."""
import ctypes

import torch

from . import _lib

F64 = torch.float64
C128 = torch.complex128


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk(t, dtype, n=None, name='tensor'):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == dtype and t.is_contiguous()):
        raise ValueError('{0}: expected a contiguous CUDA tensor of dtype {1}'.format(name, dtype))
    if n is not None and t.numel() != n:
        raise ValueError('{0}: expected {1} elements, got {2}'.format(name, n, t.numel()))
    return t.data_ptr()


def fp64_probe(mode=0, iters=4096):
    """Measured fp64 ceiling of this GPU in TFLOP/s (0: DFMA, 1: DMMA m8n8k4, 2: both interleaved)."""
    lib = _lib.load()
    _lib.init()
    tf = ctypes.c_double()
    ms = ctypes.c_double()
    _lib.check(lib.xrt_b200_fp64_probe(mode, iters, ctypes.byref(tf), ctypes.byref(ms)), 'fp64_probe')
    return tf.value, ms.value


_workspaces = {}  # device -> uint8 tensor, grown on demand and kept (packed records + split partial sums)


def kirchhoff_workspace(device, nbytes):
    """One workspace per device, reused by every call on that device's current stream (calls on one stream are
    ordered, so the next launch can overwrite it) -- r01 allocated 128 MB + 80 * M * splits bytes per call."""
    key = str(device)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        _workspaces[key] = ws = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)
    return ws


PHASE_LIMIT_TABLE_SINCOS = 2.**40 * 3.141592653589793  # sincos_tab (common.cuh) reduces with 2^11/pi: exact below this phase


def phase_bound(xM, yM, zM, pos, k):
    """A bound on |k r| over all pairs from the bounding boxes of the receiving points and of the samples (device
    reductions + one small host read): what decides whether the table-based sin / cos of the Kirchhoff kernel applies."""
    if xM.numel() == 0 or k.numel() == 0:
        return 0.
    lo = torch.stack((xM.min(), yM.min(), zM.min()))
    hi = torch.stack((xM.max(), yM.max(), zM.max()))
    p = pos[:, :3]
    plo, phi = p.min(dim=0).values, p.max(dim=0).values
    span = torch.maximum((hi - plo).abs(), (phi - lo).abs())
    return float((k.abs().max() * span.norm()).item())


def kirchhoff(xM, yM, zM, nl, Es, Ep, k, pos, nrm, out=None, convention='numpy', relaxed=False, R=0,
              workspace=None, uniform_normal=False, table_sincos='auto'):
    """a17 on device tensors.  pos, nrm: (N, 3|4) float64 row-major.  Returns (S, P, A, B, C) complex128[M].
    table_sincos: 'auto' takes the faster table-based sin / cos when phase_bound() allows it."""
    lib = _lib.load()
    M, N = xM.numel(), nl.numel()
    if table_sincos == 'auto':
        table_sincos = phase_bound(xM, yM, zM, pos, k) < PHASE_LIMIT_TABLE_SINCOS
    flags = (1 if relaxed else 0) | ((1 if convention == 'opencl' else 0) << 1) | (int(R) << 4) | \
        ((1 if uniform_normal else 0) << 9) | ((1 if table_sincos else 0) << 11)
    need = lib.xrt_kirchhoff_workspace_bytes(M, N, flags)
    if workspace is None or workspace.numel() < need:
        workspace = kirchhoff_workspace(xM.device, need)
    if out is None:
        out = torch.empty((5, M), dtype=C128, device=xM.device)
    ps, ns = pos.shape[1], nrm.shape[1]
    rc = lib.xrt_kirchhoff_f64(
        M, N, _chk(xM, F64, M, 'xM'), _chk(yM, F64, M, 'yM'), _chk(zM, F64, M, 'zM'), _chk(nl, F64, N, 'nl'),
        _chk(Es, C128, N, 'Es'), _chk(Ep, C128, N, 'Ep'), _chk(k, F64, N, 'k'), _chk(pos, F64, N * ps, 'pos'), ps,
        _chk(nrm, F64, N * ns, 'nrm'), ns, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(),
        out[3].data_ptr(), out[4].data_ptr(), flags, workspace.data_ptr(), _stream())
    _lib.check(rc, 'xrt_kirchhoff_f64')
    return out


def kirchhoff_pack(nl, Es, Ep, k, pos, nrm):
    """This rank's samples as 128-byte records (uint8 (N, 128)): what ranks all-gather when the samples are ray
    shards (waves.diffract(comm=...))."""
    lib = _lib.load()
    N = nl.numel()
    rec = torch.empty((N, 128), dtype=torch.uint8, device=nl.device)
    ps, ns = pos.shape[1], nrm.shape[1]
    rc = lib.xrt_kirchhoff_pack_f64(N, _chk(nl, F64, N, 'nl'), _chk(Es, C128, N, 'Es'), _chk(Ep, C128, N, 'Ep'),
                                    _chk(k, F64, N, 'k'), _chk(pos, F64, N * ps, 'pos'), ps,
                                    _chk(nrm, F64, N * ns, 'nrm'), ns, rec.data_ptr(), _stream())
    _lib.check(rc, 'xrt_kirchhoff_pack_f64')
    return rec


def kirchhoff_packed(xM, yM, zM, records, out=None, convention='numpy', relaxed=False, R=0, uniform_normal=False,
                     table_sincos=False):
    """a17 from packed records (kirchhoff_pack, possibly all-gathered over ranks).  table_sincos=True: the caller
    vouches for |k r| < PHASE_LIMIT_TABLE_SINCOS (phase_bound on its own samples and receiving points)."""
    lib = _lib.load()
    M, N = xM.numel(), records.shape[0]
    flags = (1 if relaxed else 0) | ((1 if convention == 'opencl' else 0) << 1) | (int(R) << 4) | \
        ((1 if uniform_normal else 0) << 9) | ((1 if table_sincos else 0) << 11)
    ws = kirchhoff_workspace(xM.device, lib.xrt_kirchhoff_workspace_bytes(M, N, flags))
    if out is None:
        out = torch.empty((5, M), dtype=C128, device=xM.device)
    if not (records.is_cuda and records.is_contiguous() and records.dtype == torch.uint8):
        raise ValueError('records: contiguous CUDA uint8 (N, 128)')
    rc = lib.xrt_kirchhoff_packed_f64(M, N, _chk(xM, F64, M, 'xM'), _chk(yM, F64, M, 'yM'), _chk(zM, F64, M, 'zM'),
                                      records.data_ptr(), out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(),
                                      out[3].data_ptr(), out[4].data_ptr(), flags, ws.data_ptr(), _stream())
    _lib.check(rc, 'xrt_kirchhoff_packed_f64')
    return out


def undulator_raw(Kx, Ky, gamma, wu, w, ww1, ddphi, ddpsi, tables, R=0, symmetric=False):
    """a4 on device tensors, the reference kernel interface: returns (Is, Ip) = Undulator._sp_sum."""
    lib = _lib.load()
    n = w.numel()
    jend = tables[0].numel()
    out = torch.empty((2, n), dtype=C128, device=w.device)
    rc = lib.xrt_undulator_f64(
        n, jend, 0.0, float(Kx), float(Ky), *[_chk(t, F64, n) for t in (gamma, wu, w, ww1, ddphi, ddpsi)],
        *[_chk(t, F64, jend) for t in tables], out[0].data_ptr(), out[1].data_ptr(),
        (int(R) << 4) | ((1 if symmetric else 0) << 10), _stream())
    _lib.check(rc, 'xrt_undulator_f64')
    return out[0], out[1]


def undulator_imap(Kx, Ky, L0, Np, gamma0, amp_const, dstep, bw_mode, w, theta, psi, tables, gamma=None, R=0,
                   out=None, symmetric=False):
    """a3+a4 fused on device tensors: returns (I, Es, Ep) of Undulator._build_I_map_conv."""
    lib = _lib.load()
    n = w.numel()
    jend = tables[0].numel()
    if out is None:
        out = (torch.empty(n, dtype=F64, device=w.device), torch.empty(n, dtype=C128, device=w.device),
               torch.empty(n, dtype=C128, device=w.device))
    I, Es, Ep = out
    rc = lib.xrt_undulator_imap_f64(
        n, jend, float(Kx), float(Ky), float(L0), float(Np), float(gamma0), float(amp_const), float(dstep),
        int(bw_mode), _chk(gamma, F64, n) if gamma is not None else None, _chk(w, F64, n), _chk(theta, F64, n),
        _chk(psi, F64, n), *[_chk(t, F64, jend) for t in tables], _chk(I, F64, n), _chk(Es, C128, n),
        _chk(Ep, C128, n), (int(R) << 4) | ((1 if symmetric else 0) << 10), _stream())
    _lib.check(rc, 'xrt_undulator_imap_f64')
    return I, Es, Ep


def undulator_ext_imap(variant, scalar0, Kx, Ky, L0, Np, gamma0, amp_const, dstep, bw_mode, w, theta, psi, tables,
                       gamma=None, out=None):
    """f3 on device tensors: (I, Es, Ep) of Undulator._build_I_map_conv for a tapered (variant 1, scalar0 =
    _taperVal) or near-field (variant 2, scalar0 = R0*pi*2/L0) undulator; Np periods summed explicitly."""
    lib = _lib.load()
    n = w.numel()
    jend = tables[0].numel()
    if out is None:
        out = (torch.empty(n, dtype=F64, device=w.device), torch.empty(n, dtype=C128, device=w.device),
               torch.empty(n, dtype=C128, device=w.device))
    I, Es, Ep = out
    rc = lib.xrt_undulator_ext_imap_f64(
        int(variant), n, jend, int(Np), float(scalar0), float(Kx), float(Ky), float(L0), float(gamma0),
        float(amp_const), float(dstep), int(bw_mode), _chk(gamma, F64, n) if gamma is not None else None,
        _chk(w, F64, n), _chk(theta, F64, n), _chk(psi, F64, n), *[_chk(t, F64, jend) for t in tables],
        _chk(I, F64, n), _chk(Es, C128, n), _chk(Ep, C128, n), _stream())
    _lib.check(rc, 'xrt_undulator_ext_imap_f64')
    return I, Es, Ep


def _amplitude_host(material, E, bIn, bOut=None, bH=None):
    import numpy as np
    lib = _lib.load()
    _lib.init()
    E = _lib.as_f64(np.atleast_1d(E))
    n = len(E)
    arrs = [E] + [None if a is None else _lib.as_f64(np.broadcast_to(np.atleast_1d(a), E.shape))
                  for a in (bIn, bOut, bH)]
    tab, tlen, offs, lens = material.host_tables()
    tab = _lib.as_f64(tab)
    mat = material.c_struct(offs, lens)
    rs = np.empty(n, dtype=np.complex128)
    rp = np.empty(n, dtype=np.complex128)
    rc = lib.xrt_get_amplitude_f64_host(n, ctypes.byref(mat), _lib.hptr(tab), tlen,
                                        *[None if a is None else _lib.hptr(a) for a in arrs],
                                        _lib.hptr(rs), _lib.hptr(rp))
    if rc == 3:
        raise ValueError(lib.xrt_b200_last_error().decode())
    _lib.check(rc, 'xrt_get_amplitude_f64_host')
    return rs, rp


def material_amplitude_host(material, E, beamInDotNormal):
    """a12: (rs, rp) of Material.get_amplitude for host arrays."""
    return _amplitude_host(material, E, beamInDotNormal)


def crystal_amplitude_host(crystal, E, beamInDotNormal, beamOutDotNormal=None, beamInDotHNormal=None):
    """a13: (curveS, curveP) of Crystal.get_amplitude (thick Bragg) for host arrays."""
    return _amplitude_host(crystal, E, beamInDotNormal, beamOutDotNormal, beamInDotHNormal)


def grid_is_symmetric(tg, ag, sintg, costg):
    """Mirror symmetry of an integration grid about tg = 0 to a few ulps (host arrays)."""
    import numpy as np
    tol = 4e-15  # np.arange / np.sin / np.cos leave asymmetries of a few ulps of pi
    return bool(np.all(np.abs(tg + tg[::-1]) <= 4 * tol) and np.all(np.abs(ag - ag[::-1]) <= tol * np.abs(ag)) and
                np.all(np.abs(sintg + sintg[::-1]) <= tol) and np.all(np.abs(costg - costg[::-1]) <= tol))
