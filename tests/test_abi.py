"""The C-ABI library loads on a CPU-only box and exports every symbol include/xrt_b200.h declares."""
import ctypes
import os
import re

import pytest

from conftest import ROOT, golden


def _declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'xrt_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(xrt_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from xrt_b200 import _lib
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), n
        assert n in _lib.PROTOTYPES, 'no ctypes prototype for ' + n


def test_struct_layouts_match_ctypes_mirrors():
    from xrt_b200 import _lib
    from xrt_b200.beams import xrt_beam
    from xrt_b200.oes import xrt_oe_params
    from xrt_b200.materials import xrt_material_params
    lib = _lib.load()
    assert lib.xrt_b200_sizeof(0) == ctypes.sizeof(xrt_beam)
    assert lib.xrt_b200_sizeof(1) == ctypes.sizeof(xrt_oe_params)
    assert lib.xrt_b200_sizeof(2) == ctypes.sizeof(xrt_material_params)
    from xrt_b200.planes import xrt_plane_params
    from xrt_b200.sources import xrt_source_params
    from xrt_b200.histogram import xrt_plot_params
    assert lib.xrt_b200_sizeof(3) == ctypes.sizeof(xrt_plane_params)
    assert lib.xrt_b200_sizeof(4) == ctypes.sizeof(xrt_source_params)
    assert lib.xrt_b200_sizeof(5) == ctypes.sizeof(xrt_plot_params)
    from xrt_b200.sources import xrt_bm_params, xrt_geom_params
    assert lib.xrt_b200_sizeof(7) == ctypes.sizeof(xrt_bm_params)
    assert lib.xrt_b200_sizeof(8) == ctypes.sizeof(xrt_geom_params)


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the product must raise, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from xrt_b200 import _lib
    from xrt_b200.cuda_backend import XRT_CUDA
    with pytest.raises(_lib.XrtB200Error):
        XRT_CUDA('diffract.cl', targetOpenCL='CUDA')


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: nothing under xrt_b200/ may import it."""
    pkg = os.path.join(ROOT, 'xrt_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', txt, flags=re.M), f


def test_constants_match_reference():
    from xrt_b200 import physconsts as pc
    from oracle import physconsts as oc
    g = golden('constants')
    for name in ('PI', 'PI2', 'SIE0', 'C', 'M0', 'EV2ERG', 'FINE_STR', 'E2WC', 'R0', 'AVOGADRO', 'CH', 'CHBAR',
                 'K2B'):
        assert getattr(pc, name) == float(g[name]), name
        assert getattr(oc, name) == float(g[name]), name
