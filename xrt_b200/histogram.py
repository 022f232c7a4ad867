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


def _state_mask(rayFlag):
    """raycing.get_output's rayFlag (raycing/__init__.py:203-218): 1 good, 2 out, 3 over, 4 alive (state > 0),
    negative = rays lost with exactly that state.  Returns (bit mask, list of negative flags)."""
    mask, dead = 0, []
    for f in rayFlag:
        f = int(f)
        if f == 4:
            mask |= 1 | 2 | 4 | 16
        elif f in (1, 2, 3):
            mask |= {1: 1, 2: 2, 3: 4}[f]
        elif f < 0:
            dead.append(f)
        else:
            raise ValueError('rayFlag {0}'.format(f))
    if len(dead) > 4:
        raise ValueError('at most 4 negative ray flags')
    return mask, dead


def hist2d(beam, xlimits, ylimits, bins=(256, 256), xattr='x', yattr='z', rayFlag=(1,), weights='intensity',
           out=None):
    """Returns (hist2d[ny, nx] CUDA tensor, yedges, xedges) == np.histogram2d(y, x, bins=[ny, nx],
    range=[ylimits, xlimits], weights=Jss + Jpp) over the rays with state in *rayFlag* (1 good, 2 out, 3 over,
    4 alive, negative: dead; as in raycing.get_output).  *out*: accumulate into an existing image (repeats)."""
    lib = _lib.load()
    nx, ny = int(bins[0]), int(bins[1])
    xe = np.linspace(xlimits[0], xlimits[1], nx + 1)  # the edges np.histogramdd builds
    ye = np.linspace(ylimits[0], ylimits[1], ny + 1)
    dev = beam.device
    dxe, dye = torch.from_numpy(xe).to(dev), torch.from_numpy(ye).to(dev)
    if out is None:
        out = torch.zeros((ny, nx), dtype=torch.float64, device=dev)
    mask, dead = _state_mask(rayFlag)
    if dead:
        mask |= 8  # this entry point takes no exact negative flags: every dead ray
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


# ---------------------------------------------------------------------------------------------- one XYCPlot
class xrt_plot_params(ctypes.Structure):
    """struct xrt_plot_params of include/xrt_b200.h"""
    _fields_ = [('x', ctypes.c_void_p), ('y', ctypes.c_void_p), ('c', ctypes.c_void_p),
                ('xfactor', ctypes.c_double), ('yfactor', ctypes.c_double), ('cfactor', ctypes.c_double),
                ('ex', ctypes.c_void_p), ('ey', ctypes.c_void_p), ('ec', ctypes.c_void_p),
                ('nx', ctypes.c_int), ('ny', ctypes.c_int), ('nc', ctypes.c_int),
                ('Jss', ctypes.c_void_p), ('Jpp', ctypes.c_void_p), ('Jsp', ctypes.c_void_p),
                ('E', ctypes.c_void_p), ('state', ctypes.c_void_p),
                ('state_mask', ctypes.c_int), ('n_dead_flags', ctypes.c_int), ('dead_flags', ctypes.c_int * 4),
                ('flux_kind', ctypes.c_int), ('srcWt', ctypes.c_double),
                ('climits', ctypes.c_double * 2), ('colorFactor', ctypes.c_double),
                ('colorSaturation', ctypes.c_double),
                ('hist2d', ctypes.c_void_p), ('histx', ctypes.c_void_p), ('histy', ctypes.c_void_p),
                ('histc', ctypes.c_void_p), ('counters', ctypes.c_void_p)]


FLUX_KINDS = (('power', 5), ('s', 1), ('p', 2), ('+/-45', 3), ('left-right', 4))


def _flux_kind(fluxKind):
    """The startswith chain of raycing.get_output (raycing/__init__.py:267-296)."""
    if fluxKind.startswith('E'):
        raise NotImplementedError('field-amplitude plots (fluxKind E*) are accumulated by xrt on the host')
    for prefix, code in FLUX_KINDS:
        if fluxKind.startswith(prefix):
            return code
    return 0


class Axis(object):
    """What multipro reads of an XYCAxis: data (Beam attribute name), factor, limits, bins."""

    def __init__(self, data, limits, bins=128, factor=1.):
        self.data, self.limits, self.bins, self.factor = data, list(limits), int(bins), float(factor)

    def edges(self):
        return np.linspace(self.limits[0], self.limits[1], self.bins + 1)


class PlotHistograms(object):
    """Device accumulators of one XYCPlot: total2D[ny, nx], total2D_RGB[ny, nx, 3], xtotal1D[nx], xtotal1D_RGB[nx, 3],
    ytotal1D..., etotal1D... (the names of xrt.plotter.XYCPlot) plus the ray counters of get_output.  `accumulate`
    adds one beam (one repeat); ray shards on several GPUs call `allreduce()` once at the end."""

    def __init__(self, xaxis, yaxis, caxis=None, fluxKind='total', rayFlag=(1,), colorFactor=1./3,
                 colorSaturation=0.85, device=None):
        self.xaxis, self.yaxis, self.caxis = xaxis, yaxis, caxis
        self.fluxKind, self.rayFlag = fluxKind, tuple(rayFlag)
        self.colorFactor, self.colorSaturation = float(colorFactor), float(colorSaturation)
        dev = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
        z = lambda *shape: torch.zeros(shape, dtype=torch.float64, device=dev)  # noqa: E731
        self._h2 = z(yaxis.bins, xaxis.bins, 4)
        self._hx = z(xaxis.bins, 4)
        self._hy = z(yaxis.bins, 4)
        self._hc = z(caxis.bins, 4) if caxis is not None else None
        self._counters = torch.zeros(6, dtype=torch.int64, device=dev)
        t = lambda a: torch.from_numpy(a.edges()).to(dev)  # noqa: E731
        self._ex, self._ey = t(xaxis), t(yaxis)
        self._ec = t(caxis) if caxis is not None else None
        self.repeats = 0

    def accumulate(self, beam):
        lib = _lib.load()
        p = xrt_plot_params()
        ptr = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
        p.x, p.y = ptr(getattr(beam, self.xaxis.data)), ptr(getattr(beam, self.yaxis.data))
        p.xfactor, p.yfactor = self.xaxis.factor, self.yaxis.factor
        p.ex, p.ey, p.nx, p.ny = ptr(self._ex), ptr(self._ey), self.xaxis.bins, self.yaxis.bins
        if self.caxis is not None:
            p.c, p.cfactor = ptr(getattr(beam, self.caxis.data)), self.caxis.factor
            p.ec, p.nc = ptr(self._ec), self.caxis.bins
            p.climits[0], p.climits[1] = self.caxis.limits
        p.Jss, p.Jpp, p.E, p.state = ptr(beam.Jss), ptr(beam.Jpp), ptr(beam.E), ptr(beam.state)
        p.Jsp = ptr(torch.view_as_real(beam.Jsp))
        p.state_mask, dead = _state_mask(self.rayFlag)
        p.n_dead_flags = len(dead)
        for k, v in enumerate(dead):
            p.dead_flags[k] = v
        p.flux_kind = _flux_kind(self.fluxKind)
        p.srcWt = len(beam) * beam.sourceWeight if hasattr(beam, 'sourceWeight') else 1.
        p.colorFactor, p.colorSaturation = self.colorFactor, self.colorSaturation
        p.hist2d, p.histx, p.histy, p.histc = ptr(self._h2), ptr(self._hx), ptr(self._hy), ptr(self._hc)
        p.counters = ptr(self._counters)
        rc = lib.xrt_plot_hist_f64(len(beam), ctypes.byref(p),
                                   ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, 'xrt_plot_hist_f64')
        self.repeats += 1
        return self

    def allreduce(self):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            for t in (self._h2, self._hx, self._hy, self._hc, self._counters):
                if t is not None:
                    dist.all_reduce(t)
        return self

    # ---- views with the reference's names
    total2D = property(lambda self: self._h2[..., 0])
    total2D_RGB = property(lambda self: self._h2[..., 1:])
    xtotal1D = property(lambda self: self._hx[:, 0])
    xtotal1D_RGB = property(lambda self: self._hx[:, 1:])
    ytotal1D = property(lambda self: self._hy[:, 0])
    ytotal1D_RGB = property(lambda self: self._hy[:, 1:])
    etotal1D = property(lambda self: self._hc[:, 0])
    etotal1D_RGB = property(lambda self: self._hc[:, 1:])

    def counters(self):
        """dict(nraysSelected, nraysAlive, nraysGood, nraysOut, nraysOver, nraysDead) -- get_output's counts"""
        v = self._counters.tolist()
        return dict(zip(('nraysSelected', 'nraysAlive', 'nraysGood', 'nraysOut', 'nraysOver', 'nraysDead'), v))

    def to_numpy(self):
        out = {k: getattr(self, k).cpu().numpy() for k in
               ('total2D', 'total2D_RGB', 'xtotal1D', 'xtotal1D_RGB', 'ytotal1D', 'ytotal1D_RGB')}
        if self._hc is not None:
            out['etotal1D'] = self.etotal1D.cpu().numpy()
            out['etotal1D_RGB'] = self.etotal1D_RGB.cpu().numpy()
        out.update(self.counters())
        return out
