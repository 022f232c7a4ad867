"""Oracle: Fresnel-Kirchhoff diffraction integral (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Follows waves._diffraction_integral_conv (xrt/backends/raycing/waves.py:834-851) operation by operation; the
only change is that the M x N matrices are built in row blocks so that memory stays bounded.
"""
import numpy as np

from .physconsts import CHBAR


def diffraction_integral(x, y, z, n, nl, Es, Ep, E, xDiffr, yDiffr, zDiffr, block=256):
    """x, y, z, nl, Es, Ep, E: the *good* input samples (already masked, waves.py:836-838 `[good]`);
    n: 3-sequence of scalars or arrays (surface normal at the samples); xDiffr..: the M receiving points.
    Returns Es, Ep, aE, bE, cE (complex128[M]) exactly as waves.py:851."""
    x = np.asarray(x, dtype=np.float64)
    M = len(xDiffr)
    k = E / CHBAR * 1e7  # [mm^-1]  waves.py:842
    EsP = Es + Ep
    outs = [np.zeros(M, dtype=np.complex128) for _ in range(5)]
    for i0 in range(0, M, block):
        sl = slice(i0, min(M, i0 + block))
        a = xDiffr[sl, np.newaxis] - x  # waves.py:836-838
        b = yDiffr[sl, np.newaxis] - y
        c = zDiffr[sl, np.newaxis] - z
        pathAfter = (a**2 + b**2 + c**2)**0.5  # waves.py:839
        ns = (a*n[0] + b*n[1] + c*n[2]) / pathAfter  # waves.py:840
        U = k*1j/(4*np.pi) * (nl+ns) * np.exp(1j*k*(pathAfter)) / pathAfter  # waves.py:844
        outs[0][sl] = (Es * U).sum(axis=1)  # waves.py:845-846
        outs[1][sl] = (Ep * U).sum(axis=1)
        abcU = k**2/(4*np.pi) * EsP * U / pathAfter  # waves.py:847
        outs[2][sl] = (abcU * a).sum(axis=1)  # waves.py:848-850
        outs[3][sl] = (abcU * b).sum(axis=1)
        outs[4][sl] = (abcU * c).sum(axis=1)
    return tuple(outs)


def term_magnitude_sum(x, y, z, n, nl, Es, Ep, E, xDiffr, yDiffr, zDiffr, block=256):
    """sum_j |term_ij| for the five outputs: the scale against which parity tolerances are stated
    ('rel 1e-11 * sum|terms|')."""
    M = len(xDiffr)
    k = E / CHBAR * 1e7
    outs = [np.zeros(M) for _ in range(5)]
    for i0 in range(0, M, block):
        sl = slice(i0, min(M, i0 + block))
        a = xDiffr[sl, np.newaxis] - x
        b = yDiffr[sl, np.newaxis] - y
        c = zDiffr[sl, np.newaxis] - z
        r = (a**2 + b**2 + c**2)**0.5
        ns = (a*n[0] + b*n[1] + c*n[2]) / r
        U = np.abs(k/(4*np.pi) * (nl+ns) / r)
        outs[0][sl] = (np.abs(Es) * U).sum(axis=1)
        outs[1][sl] = (np.abs(Ep) * U).sum(axis=1)
        abcU = k**2/(4*np.pi) * np.abs(Es+Ep) * U / r
        outs[2][sl] = (abcU * np.abs(a)).sum(axis=1)
        outs[3][sl] = (abcU * np.abs(b)).sum(axis=1)
        outs[4][sl] = (abcU * np.abs(c)).sum(axis=1)
    return tuple(outs)
