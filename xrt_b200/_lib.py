# -*- coding: utf-8 -*-
"""ctypes binding of libxrt_b200.so (C ABI declared in include/xrt_b200.h).

The library is built in-tree by `__graft_entry__.build()` (nvcc, sm_100a).  There is no CPU fallback: if the
shared object is missing, or no CUDA device is present when a compute entry point is called, an exception is
raised (north star: "no Triton, no multi-backend dispatch, no CPU fallback")."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# XRT_B200_LIB names another build of the same library (kernel-variant sweeps, tools/build_variant.sh)
LIB_PATH = os.environ.get('XRT_B200_LIB') or os.path.join(_HERE, 'libxrt_b200.so')

c_dp = ctypes.c_void_p  # device or host pointer passed as an integer address
c_sz = ctypes.c_size_t
c_int = ctypes.c_int
c_dbl = ctypes.c_double

# name -> (restype, argtypes); must list every symbol include/xrt_b200.h declares (tests check this).
PROTOTYPES = {
    'xrt_b200_version': (c_int, []),
    'xrt_b200_init': (c_int, [c_int]),
    'xrt_b200_device_count': (c_int, []),
    'xrt_b200_sm_count': (c_int, []),
    'xrt_b200_last_error': (ctypes.c_char_p, []),
    'xrt_b200_sync': (c_int, [c_dp]),
    'xrt_b200_sizeof': (c_sz, [c_int]),
    'xrt_b200_fp64_probe': (c_int, [c_int, c_int, ctypes.POINTER(c_dbl), ctypes.POINTER(c_dbl)]),
    'xrt_b200_math_probe_host': (c_int, [c_int, c_sz, c_dp, c_dp, c_dp]),
    'xrt_kirchhoff_workspace_bytes': (c_sz, [c_sz, c_sz, c_int]),
    'xrt_kirchhoff_f64': (c_int, [c_sz, c_sz] + [c_dp] * 8 + [c_int, c_dp, c_int] + [c_dp] * 5 +
                          [c_int, c_dp, c_dp]),
    'xrt_kirchhoff_pack_f64': (c_int, [c_sz] + [c_dp] * 5 + [c_int, c_dp, c_int, c_dp, c_dp]),
    'xrt_kirchhoff_packed_f64': (c_int, [c_sz, c_sz] + [c_dp] * 9 + [c_int, c_dp, c_dp]),
    'xrt_kirchhoff_f64_host': (c_int, [c_sz, c_sz] + [c_dp] * 8 + [c_int, c_dp, c_int] + [c_dp] * 5 + [c_int]),
    'xrt_undulator_f64': (c_int, [c_sz, c_int, c_dbl, c_dbl, c_dbl] + [c_dp] * 14 + [c_int, c_dp]),
    'xrt_undulator_f64_host': (c_int, [c_sz, c_int, c_dbl, c_dbl, c_dbl] + [c_dp] * 14 + [c_int]),
    'xrt_reflect_f64': (c_int, [c_sz, c_dp, c_dp, c_dp, c_sz, c_dp, c_dp, c_dp, c_dp, c_dp, c_int,
                                ctypes.c_uint64, c_dp]),
    'xrt_reflect_expose_f64': (c_int, [c_sz, c_dp, c_dp, c_dp, c_sz, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_int,
                                       ctypes.c_uint64, c_dp]),
    'xrt_reflect_multi_deriv_f64': (c_int, [c_sz, c_dp, c_dp, c_dp, c_dp, c_int, c_dp]),
    'xrt_reflect_multi_elevation_f64': (c_int, [c_sz, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]),
    'xrt_reduction_decode': (c_dbl, [ctypes.c_uint64]),
    'xrt_expose_f64': (c_int, [c_sz, c_dp, c_dp, c_dp, c_dp, c_dp]),
    'xrt_propagate_rect_f64': (c_int, [c_sz, c_dp, c_dp, c_dp, c_dp, c_dp]),
    'xrt_expose_hemispheric_f64': (c_int, [c_sz, c_dp, c_dbl, c_dbl, c_dbl, c_dp, c_dp, c_dp]),
    'xrt_get_amplitude_f64_host': (c_int, [c_sz, c_dp, c_dp, c_sz] + [c_dp] * 6),
    'xrt_source_sample_f64': (c_int, [c_sz, ctypes.c_uint64, ctypes.c_uint64] + [c_dp] * 10 + [c_dp]),
    'xrt_source_fill_f64': (c_int, [c_sz, c_dp, c_dp] + [c_dp] * 9 + [c_dp, c_sz, c_dp]),
    'xrt_shine_uniform_f64': (c_int, [c_sz, ctypes.c_uint64, ctypes.c_uint64, c_dp, c_int] + [c_dbl] * 7 +
                              [c_int, c_int] + [c_dp] * 6 + [c_dp, c_dp, c_dp, c_dp]),
    'xrt_oe_surface_f64': (c_int, [c_sz, c_dp] + [c_dp] * 3 + [c_dp]),
    'xrt_oe_normals_f64': (c_int, [c_sz, c_dp] + [c_dp] * 6 + [c_dp]),
    'xrt_multilayer_f64': (c_int, [c_sz, c_int] + [c_dp] * 12 + [c_dp]),
    'xrt_multilayer_f64_host': (c_int, [c_sz, c_int] + [c_dp] * 12),
    'xrt_multilayer_tran_f64': (c_int, [c_sz, c_int, c_dbl, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]),
    'xrt_multilayer_tran_f64_host': (c_int, [c_sz, c_int, c_dbl, c_dp, c_dp, c_dp, c_dp, c_dp]),
    'xrt_plot_hist_f64': (c_int, [c_sz, c_dp, c_dp]),
    'xrt_hist2d_f64': (c_int, [c_sz, c_dp, c_dp, c_dp, c_dp, c_dp, c_int, c_dp, c_int, c_dp, c_int, c_dp, c_dp]),
    'xrt_undulator_ext_f64': (c_int, [c_int, c_sz, c_int, c_int, c_dbl, c_dbl, c_dbl] + [c_dp] * 14 + [c_dp]),
    'xrt_undulator_ext_f64_host': (c_int, [c_int, c_sz, c_int, c_int, c_dbl, c_dbl, c_dbl] + [c_dp] * 14),
    'xrt_undulator_ext_imap_f64': (c_int, [c_int, c_sz, c_int, c_int] + [c_dbl] * 7 + [c_int] + [c_dp] * 13 +
                                   [c_dp]),
    'xrt_geometric_source_f64': (c_int, [c_sz, c_dp, ctypes.c_uint64, ctypes.c_uint64, c_dp, c_dp, c_dp, c_int, c_dp]),
    'xrt_bm_imap_f64': (c_int, [c_sz, c_dp] + [c_dp] * 7 + [c_dp]),
    'xrt_bm_imap_f64_host': (c_int, [c_sz, c_dp] + [c_dp] * 7),
    'xrt_bm_fill_f64': (c_int, [c_sz, c_dp, c_dp, c_dp] + [c_dp] * 11 + [c_dp, c_sz, c_dp]),
    'xrt_bm_shine_uniform_f64': (c_int, [c_sz, ctypes.c_uint64, ctypes.c_uint64, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]),
    'xrt_custom_field_f64': (c_int, [c_sz, c_int, c_dbl, c_dbl] + [c_dp] * 16 + [c_dp]),
    'xrt_custom_field_f64_host': (c_int, [c_sz, c_int, c_dbl, c_dbl] + [c_dp] * 16),
    'xrt_compact_scratch_words': (c_sz, [c_sz]),
    'xrt_compact_rays_f64': (c_int, [c_sz, c_dp, c_int, c_int, c_dp, c_dp, c_dp, c_dp, c_dp]),
    'xrt_nccl_unique_id': (c_int, [c_dp]),
    'xrt_nccl_comm_init': (c_int, [c_int, c_int, c_dp, ctypes.POINTER(ctypes.c_void_p)]),
    'xrt_nccl_comm_destroy': (c_int, [c_dp]),
    'xrt_allreduce_slots': (c_int, [c_dp, c_int, c_dp, c_int, c_dp, c_dp]),
    'xrt_slot_kind': (c_int, [c_int]),
    'xrt_allgather_beam': (c_int, [c_dp, c_int, c_int, c_dp, c_dp, c_dp, c_dp]),
    'xrt_allgather_bytes': (c_int, [c_dp, c_int, c_int, c_dp, c_dp, c_dp, c_dp]),
    'xrt_undulator_imap_f64': (c_int, [c_sz, c_int] + [c_dbl] * 7 + [c_int] + [c_dp] * 13 + [c_int, c_dp]),
}


class XrtB200Error(RuntimeError):
    pass


_lib = None
_initialised_device = None


def load():
    """Load the shared library (does not touch the GPU)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise XrtB200Error(
                "{0} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` at the repo root. "
                "xrt_b200 has no CPU fallback.".format(LIB_PATH))
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc, what=''):
    if rc != 0:
        msg = load().xrt_b200_last_error().decode('utf-8', 'replace')
        raise XrtB200Error('{0} failed (rc={1}): {2}'.format(what or 'xrt_b200 call', rc, msg))


def init(device=None):
    """Bind this process to one GPU (one process per GPU; LOCAL_RANK picks it under torchrun)."""
    global _initialised_device
    lib = load()
    if device is None:
        device = int(os.environ.get('LOCAL_RANK', '0'))
    if _initialised_device != device:
        check(lib.xrt_b200_init(int(device)), 'xrt_b200_init')
        _initialised_device = device
    return device


def hptr(arr):
    """Address of a C-contiguous numpy array (host pointer)."""
    assert arr.flags['C_CONTIGUOUS'] or arr.flags['F_CONTIGUOUS']
    return arr.ctypes.data


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def as_c128(a):
    return np.ascontiguousarray(a, dtype=np.complex128)
