# -*- coding: utf-8 -*-
"""Device histogramming of beams: the weighted 2-D histogram xrt's plot layer accumulates
per repeat -- raycing.get_output (ray selection by state, intensity = Jss + Jpp; raycing/__init__.py:170-300) +
do_hist2d (xrt/multipro.py:119-171) -- so that 1e8..1e9-ray beams never have to leave the GPU: only the image
(ny x nx doubles) crosses PCIe, and ray shards on several GPUs are combined with one all-reduce of the image."""
import ctypes

import numpy as np
import torch

from . import _lib

STATE_BITS = {1: 1, 2: 2, 3: 4, 'dead': 8}


def hist2d(beam, xlimits, ylimits, bins=(256, 256), xattr='x', yattr='z', rayFlag=(1,), weights='intensity',
           out=None):
    """Returns (hist2d[ny, nx] CUDA tensor, yedges, xedges) == np.histogram2d(y, x, bins=[ny, nx],
    range=[ylimits, xlimits], weights=Jss + Jpp) over the rays with state in *rayFlag* (1 good, 2 out, 3 over,
    4 dead, as in raycing.get_output).  *out*: accumulate into an existing image (repeats)."""
    lib = _lib.load()
    nx, ny = int(bins[0]), int(bins[1])
    xe = np.linspace(xlimits[0], xlimits[1], nx + 1)  # the edges np.histogramdd builds
    ye = np.linspace(ylimits[0], ylimits[1], ny + 1)
    dev = beam.device
    dxe, dye = torch.from_numpy(xe).to(dev), torch.from_numpy(ye).to(dev)
    if out is None:
        out = torch.zeros((ny, nx), dtype=torch.float64, device=dev)
    mask = 0
    for f in rayFlag:
        mask |= {1: 1, 2: 2, 3: 4, 4: 8}[int(f)]
    x, y = getattr(beam, xattr), getattr(beam, yattr)
    w1 = w2 = None
    if weights == 'intensity':
        w1, w2 = beam.Jss, beam.Jpp
    elif weights is not None:
        w1 = weights
    ptr = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
    rc = lib.xrt_hist2d_f64(len(beam), ptr(x), ptr(y), ptr(w1), ptr(w2), beam.state.data_ptr(), mask, dxe.data_ptr(),
                            nx, dye.data_ptr(), ny, out.data_ptr(),
                            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, 'xrt_hist2d_f64')
    return out, ye, xe


def allreduce_image(image):
    """Sum the images of all ray shards (one NCCL all-reduce of ny x nx doubles)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(image)
    return image
