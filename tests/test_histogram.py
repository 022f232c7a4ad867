"""f1 (next row): device 2-D histogram == np.histogram2d as xrt's do_hist2d calls it (xrt/multipro.py:119-171)."""
import numpy as np
import pytest

from conftest import golden
from helpers import oracle_beam, device_beam


@pytest.mark.gpu
@pytest.mark.parametrize('bins', [(256, 256), (64, 33), (1, 1)])
def test_cuda_hist2d_matches_numpy(bins, cuda_device):
    from xrt_b200 import histogram
    g = golden('expose_screen')
    b = oracle_beam(g, 'out_')
    # put a few rays exactly on edges / outside the range / NaN to exercise the edge rules of np.histogramdd
    xl, zl = [-1.0, 1.5], [-0.8, 0.9]
    b.x[:5] = [xl[0], xl[1], np.nextafter(xl[1], 2), xl[0] - 1e-9, np.nan]
    b.z[5:8] = [zl[0], zl[1], np.nextafter(zl[0], -1)]
    b.state[:8] = 1
    good = b.state == 1
    ref, ye, xe = np.histogram2d(b.z[good], b.x[good], bins=[bins[1], bins[0]], range=[zl, xl],
                                 weights=(b.Jss + b.Jpp)[good])
    img, ye2, xe2 = histogram.hist2d(device_beam(b), xl, zl, bins=bins)
    assert np.array_equal(ye, ye2) and np.array_equal(xe, xe2)
    out = img.cpu().numpy()
    assert out.shape == ref.shape
    # identical bin membership; sums differ only by the order of the additions
    assert np.array_equal(out != 0, ref != 0)
    assert np.abs(out - ref).max() <= 1e-12 * ref.max()
    # unweighted counts are exact
    cnt, _, _ = histogram.hist2d(device_beam(b), xl, zl, bins=bins, weights=None, rayFlag=(4, -1))  # alive + every dead ray
    refc, _, _ = np.histogram2d(b.z, b.x, bins=[bins[1], bins[0]], range=[zl, xl])
    assert np.array_equal(cnt.cpu().numpy(), refc)
    # accumulation over repeats
    img2, _, _ = histogram.hist2d(device_beam(b), xl, zl, bins=bins, out=img)
    assert img2 is img and np.abs(img.cpu().numpy() - 2 * ref).max() <= 2e-12 * ref.max()


CASE_AXES = dict(x=('x', 'xlimits', 0, 'xfactor'), y=('z', 'ylimits', 1, 'yfactor'), c=('E', 'climits', 2, 'cfactor'))


def _axes(g, mod):
    ax = {}
    for k, (attr, lim, ib, fac) in CASE_AXES.items():
        f = float(g[fac])
        ax[k] = mod.Axis(attr, [float(g[lim][0]) * f, float(g[lim][1]) * f], int(g['bins'][ib]), f)
    return ax


def _cases(g):
    for n, (fk, fl) in enumerate(zip(g['fluxKinds'], g['rayFlags'])):
        yield n, str(fk), tuple(int(v) for v in str(fl).split(','))


def test_oracle_get_output_and_histograms_bit_exact():
    """oracle get_output == the reference's raycing.get_output on the same beam, for every flux kind / flag set."""
    from oracle import histogram as oh
    g = golden('plot_histograms')
    b = oracle_beam(g, 'b_')
    b.sourceWeight = float(g['sourceWeight'])
    ax = _axes(g, oh)
    for n, fk, flags in _cases(g):
        sel = oh.get_output(b, ax['x'], ax['y'], ax['c'], fk, flags)
        for k, v in zip(('x', 'y', 'intensity', 'flux', 'cData'), sel[:5]):
            assert np.array_equal(v, g['c%d_%s' % (n, k)]), (fk, k)
        assert [int(v) for v in sel[5:11]] == list(g['c%d_counts' % n]), fk
        h = oh.plot_histograms(b, ax['x'], ax['y'], ax['c'], fk, flags)
        for k in ('total2D', 'total2D_RGB', 'xtotal1D', 'xtotal1D_RGB', 'ytotal1D', 'ytotal1D_RGB', 'etotal1D',
                  'etotal1D_RGB'):
            assert np.array_equal(h[k], g['c%d_%s' % (n, k)]), (fk, k)


@pytest.mark.gpu
def test_cuda_plot_histograms_vs_golden(cuda_device):
    from xrt_b200 import histogram
    g = golden('plot_histograms')
    b = oracle_beam(g, 'b_')
    db = device_beam(b)
    db.sourceWeight = float(g['sourceWeight'])
    ax = _axes(g, histogram)
    for n, fk, flags in _cases(g):
        ph = histogram.PlotHistograms(ax['x'], ax['y'], ax['c'], fluxKind=fk, rayFlag=flags)
        ph.accumulate(db)
        out = ph.to_numpy()
        cnt = list(g['c%d_counts' % n])  # nrays, alive, good, out, over, dead
        assert out['nraysSelected'] == len(g['c%d_x' % n]), fk
        assert [out[k] for k in ('nraysAlive', 'nraysGood', 'nraysOut', 'nraysOver', 'nraysDead')] == cnt[1:], fk
        for k in ('total2D', 'total2D_RGB', 'xtotal1D', 'xtotal1D_RGB', 'ytotal1D', 'ytotal1D_RGB', 'etotal1D',
                  'etotal1D_RGB'):
            ref = g['c%d_%s' % (n, k)]
            assert out[k].shape == ref.shape, (fk, k)
            # signed weights (+/-45, left-right) cancel inside a bin: judge against the largest |bin| of the plot
            assert np.abs(out[k] - ref).max() <= 1e-12 * np.abs(ref).max(), (fk, k)
        # a second repeat doubles everything
        ph.accumulate(db)
        assert ph.repeats == 2
        ref = g['c%d_total2D' % n]
        assert np.abs(ph.total2D.cpu().numpy() - 2 * ref).max() <= 2e-12 * np.abs(ref).max()
