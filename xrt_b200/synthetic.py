# -*- coding: utf-8 -*-
"""Seeded synthetic workloads of the configurations in BASELINE.json (used by bench.py, smoke() and the
parity tests; no dependency on the oracle)."""
import numpy as np

from .physconsts import CHBAR


def kirchhoff_case(seed, N, M, E0=9000., dist=1e4, halfIn=0.05, halfOut=0.5, mesh=False, tilt=False):
    """Wave config (BASELINE.json configs[3]): N samples on a slit of +-halfIn mm, M receiving points on a
    screen `dist` mm downstream.  Returns a dict in the layout of waves._diffraction_integral_CL
    (xrt/backends/raycing/waves.py:864-890)."""
    rng = np.random.RandomState(seed)
    x = rng.uniform(-halfIn, halfIn, N)
    z = rng.uniform(-halfIn, halfIn, N)
    y = rng.normal(0, 1e-3, N) if tilt else np.zeros(N)
    E = E0 * (1 + 1e-4 * rng.uniform(-1, 1, N))
    Es = rng.normal(size=N) + 1j * rng.normal(size=N)
    Ep = 0.3 * (rng.normal(size=N) + 1j * rng.normal(size=N))
    a = rng.normal(0, 1e-4, N)
    c = rng.normal(0, 1e-4, N)
    b = np.sqrt(1 - a**2 - c**2)
    if tilt:
        n = rng.normal(size=(3, N)) * 0.05 + np.array([[0.], [1.], [0.]])
        n /= np.linalg.norm(n, axis=0)
    else:
        n = np.array([[0.], [1.], [0.]]) * np.ones((3, N))
    nl = a * n[0] + b * n[1] + c * n[2]
    if mesh:
        side = int(round(M ** 0.5))
        assert side * side == M
        gx, gz = np.meshgrid(np.linspace(-halfOut, halfOut, side), np.linspace(-halfOut, halfOut, side))
        xM, zM = gx.ravel().copy(), gz.ravel().copy()
    else:
        xM = rng.uniform(-halfOut, halfOut, M)
        zM = rng.uniform(-halfOut, halfOut, M)
    yM = dist + (rng.uniform(-1, 1, M) if tilt else np.zeros(M))
    k = E / CHBAR * 1e7
    pos = np.array([x, y, z, 0 * z], order='F')
    nrm = np.array([n[0], n[1], n[2], 0 * n[2]], order='F')
    return dict(N=N, M=M, x=x, y=y, z=z, E=E, Es=Es, Ep=Ep, n=n, nl=nl, k=k, pos=pos, nrm=nrm,
                xM=xM, yM=yM, zM=zM)
