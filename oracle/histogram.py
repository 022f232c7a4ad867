"""Oracle: what one XYCPlot accumulates per repeat (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Restates in numpy
  * raycing.get_output               xrt/backends/raycing/__init__.py:170-300  (ray selection, flux kinds, counters)
  * the body of multipro's run()     xrt/multipro.py:316-340                   (cData01, HSV -> RGB)
  * do_hist1d / do_hist2d            xrt/multipro.py:53-171                    (np.histogram / np.histogram2d)
matplotlib.colors.hsv_to_rgb is a third-party function absent from this image and from /root/reference (xrt
declares matplotlib unpinned); its published sextant algorithm is restated in hsv_to_rgb() below.  get_output is
pinned against the reference's own function on seeded beams (tools/gen_golden.py `hist`); the histogram calls are
numpy's own."""
import numpy as np

SIE0 = 1.602176565e-19


class Axis(object):
    def __init__(self, data, limits, bins=128, factor=1.):
        self.data, self.limits, self.bins, self.factor = data, list(limits), int(bins), float(factor)


def get_output(beam, xaxis, yaxis, caxis, fluxKind='total', rayFlag=(1,)):
    beamState = beam.state
    nrays = len(beam.x)
    locAlive = (beamState > 0).sum()
    part = np.zeros(nrays, dtype=bool)
    locGood = locOut = locOver = locDead = 0
    for flag in rayFlag:
        locPart = beamState > 0 if flag == 4 else beamState == flag
        if flag == 1:
            locGood = locPart.sum()
        if flag == 2:
            locOut = locPart.sum()
        if flag == 3:
            locOver = locPart.sum()
        if flag < 0:
            locDead += locPart.sum()
        part = part | locPart
    x = getattr(beam, xaxis.data) * xaxis.factor
    y = getattr(beam, yaxis.data) * yaxis.factor
    cData = getattr(beam, caxis.data) * caxis.factor
    srcWt = nrays * beam.sourceWeight if hasattr(beam, 'sourceWeight') else 1.
    if fluxKind.startswith('power'):
        intensity = (beam.Jss + beam.Jpp) * beam.E * SIE0 * srcWt
    elif fluxKind.startswith('s'):
        intensity = beam.Jss * srcWt
    elif fluxKind.startswith('p'):
        intensity = beam.Jpp * srcWt
    elif fluxKind.startswith('+/-45'):
        intensity = 2*beam.Jsp.real * srcWt
    elif fluxKind.startswith('left-right'):
        intensity = 2*beam.Jsp.imag * srcWt
    else:
        intensity = (beam.Jss + beam.Jpp) * srcWt
    flux = intensity
    return (x[part], y[part], intensity[part], flux[part], cData[part], nrays, locAlive, locGood, locOut, locOver,
            locDead)


def hsv_to_rgb(hsv):
    """matplotlib.colors.hsv_to_rgb on an (..., 3) array: i = int(6h), f = 6h - i, p = v(1-s), q = v(1-sf),
    t = v(1-s(1-f)); (r, g, b) by sextant i % 6; s == 0 -> grey."""
    hsv = np.asarray(hsv, dtype=float)
    h, s, v = hsv[..., 0], hsv[..., 1], hsv[..., 2]
    r, g, b = np.empty_like(h), np.empty_like(h), np.empty_like(h)
    i = (h * 6.0).astype(int)
    f = (h * 6.0) - i
    p = v * (1.0 - s)
    q = v * (1.0 - (s * f))
    t = v * (1.0 - (s * (1.0 - f)))
    for k, (rr, gg, bb) in enumerate(((v, t, p), (q, v, p), (p, v, t), (p, q, v), (t, p, v), (v, p, q))):
        idx = i % 6 == k
        r[idx], g[idx], b[idx] = rr[idx], gg[idx], bb[idx]
    idx = s == 0
    r[idx], g[idx], b[idx] = v[idx], v[idx], v[idx]
    return np.stack([r, g, b], axis=-1)


def colour_weights(cData, flux, climits, colorFactor, colorSaturation):
    """multipro.py:316-325"""
    cData01 = ((cData - climits[0]) * colorFactor / (climits[1] - climits[0])).reshape(-1, 1)
    cData01[cData01 < 0] = 0.
    cData01[cData01 > 1] = 1.
    cDataHSV = np.dstack((cData01, np.ones_like(cData01) * colorSaturation, flux.reshape(-1, 1)))
    return hsv_to_rgb(cDataHSV).reshape(-1, 3)


def do_hist1d(x, intensity, cDataRGB, axis):
    hist1dRGB = np.zeros((axis.bins, 3))
    hist1d, binEdges = np.histogram(x, bins=axis.bins, range=axis.limits, weights=intensity)
    for i in range(3):
        hist1dRGB[:, i], binEdges = np.histogram(x, bins=axis.bins, range=axis.limits, weights=cDataRGB[:, i])
    return hist1d, hist1dRGB, binEdges


def do_hist2d(x, y, intensity, cDataRGB, xaxis, yaxis):
    xyrange = [yaxis.limits, xaxis.limits]
    xybins = [yaxis.bins, xaxis.bins]
    hist2d, yedges, xedges = np.histogram2d(y, x, bins=xybins, range=xyrange, weights=intensity)
    hist2dRGB = np.zeros((xybins[0], xybins[1], 3))
    if len(x) > 0:
        for i in range(3):
            hist2dRGB[:, :, i], yedges, xedges = np.histogram2d(y, x, bins=xybins, range=xyrange,
                                                                weights=cDataRGB[:, i])
    return hist2d, hist2dRGB


def plot_histograms(beam, xaxis, yaxis, caxis, fluxKind='total', rayFlag=(1,), colorFactor=1./3, colorSaturation=0.85,
                    selected=None):
    """One repeat of one plot.  *selected*: the tuple get_output returned (the reference's, when pinning)."""
    if selected is None:
        selected = get_output(beam, xaxis, yaxis, caxis, fluxKind, rayFlag)
    x, y, intensity, flux, cData = selected[:5]
    rgb = colour_weights(cData, flux, caxis.limits, colorFactor, colorSaturation)
    out = {}
    out['xtotal1D'], out['xtotal1D_RGB'], _ = do_hist1d(x, flux, rgb, xaxis)
    out['ytotal1D'], out['ytotal1D_RGB'], _ = do_hist1d(y, flux, rgb, yaxis)
    out['etotal1D'], out['etotal1D_RGB'], _ = do_hist1d(cData, flux, rgb, caxis)
    out['total2D'], out['total2D_RGB'] = do_hist2d(x, y, intensity, rgb, xaxis, yaxis)
    out.update(nraysSelected=len(x), nraysAlive=int(selected[6]), nraysGood=int(selected[7]),
               nraysOut=int(selected[8]), nraysOver=int(selected[9]), nraysDead=int(selected[10]))
    return out
