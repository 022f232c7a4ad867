# -*- coding: utf-8 -*-
"""xrt_b200 -- B200-native (sm_100a) hot path of xrt's `raycing` backend behind xrt's own accelerator API.

Boundary: `XRT_CUDA` mirrors `XRT_CL` (xrt/backends/raycing/myopencl.py:86-583) and dispatches
`run_parallel(kernelName, ...)` to hand-written CUDA kernels through the C ABI in include/xrt_b200.h."""
from . import _lib  # noqa: F401
from ._lib import XrtB200Error  # noqa: F401

__version__ = '0.1.0'
