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


def geometric_beam(seed, nrays, dx=0.1, dz=0.05, dxprime=1e-4, dzprime=5e-5, energies=(8990., 9010.),
                   azimuth=0., withAmplitudes=True, kill=0.02, center=(0., 0., 0.)):
    """Seeded stand-in for GeometricSource.shine (config 1/3 input): normal position/divergence, flat energy
    band, generic complex polarisation state, a few dead rays.  Returns a dict of numpy arrays."""
    rng = np.random.RandomState(seed)
    x = rng.normal(0, dx, nrays)
    z = rng.normal(0, dz, nrays)
    y = np.zeros(nrays)
    a = rng.normal(0, dxprime, nrays)
    c = rng.normal(0, dzprime, nrays)
    b = np.ones(nrays)
    norm = (a**2 + b**2 + c**2)**0.5
    a, b, c = a / norm, b / norm, c / norm
    sa, ca = np.sin(azimuth), np.cos(azimuth)
    x, y = ca * x + sa * y, -sa * x + ca * y
    a, b = ca * a + sa * b, -sa * a + ca * b
    E = rng.uniform(energies[0], energies[1], nrays)
    Es = np.exp(1j * rng.uniform(0, 2 * np.pi, nrays))
    Ep = 0.2 * np.exp(1j * rng.uniform(0, 2 * np.pi, nrays))
    state = np.ones(nrays, dtype=np.int32)
    state[rng.rand(nrays) < kill] = -7
    out = dict(x=x + center[0], y=y + center[1], z=z + center[2], a=a, b=b, c=c, path=np.zeros(nrays), E=E,
               Jss=np.abs(Es)**2, Jpp=np.abs(Ep)**2, Jsp=Es * np.conj(Ep), state=state)
    if withAmplitudes:
        out['Es'] = Es
        out['Ep'] = Ep
    return out
