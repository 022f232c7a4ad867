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
    cnt, _, _ = histogram.hist2d(device_beam(b), xl, zl, bins=bins, weights=None, rayFlag=(1, 2, 3, 4))
    refc, _, _ = np.histogram2d(b.z, b.x, bins=[bins[1], bins[0]], range=[zl, xl])
    assert np.array_equal(cnt.cpu().numpy(), refc)
    # accumulation over repeats
    img2, _, _ = histogram.hist2d(device_beam(b), xl, zl, bins=bins, out=img)
    assert img2 is img and np.abs(img.cpu().numpy() - 2 * ref).max() <= 2e-12 * ref.max()
