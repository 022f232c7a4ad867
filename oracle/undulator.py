"""Oracle: undulator far-field integral (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Restates, in numpy and in the reference's operation order:
  * IntegratedSource._clenshaw_curtis        xrt/backends/raycing/sources/sybase.py:1106-1139
  * Undulator._build_integration_grid        xrt/backends/raycing/sources/synchr.py:1787-1801
  * Undulator._sp_sum (far field, no taper)  synchr.py:1929-2033 (the `else` branch :1997-2003 + :2005-2031)
  * Undulator._build_I_map_conv              synchr.py:2048-2108
"""
import numpy as np

from .physconsts import PI, PI2, E2WC, FINE_STR, SIE0, EV2ERG, M0, C


def gamma_from_eE(eE_GeV):
    return float(eE_GeV) * 1e9 * EV2ERG / (M0 * C**2)  # sybase.py:155


def clenshaw_curtis(n):
    points = -np.cos((np.pi * np.arange(n)) / (n - 1))
    if n == 2:
        return points, np.array([1.0, 1.0])
    n -= 1
    N = np.arange(1, n, 2)
    length = len(N)
    m = n - length
    v0 = np.concatenate([2.0 / N / (N - 2), np.array([1.0 / N[-1]]), np.zeros(m)])
    v2 = -v0[:-1] - v0[:0:-1]
    g0 = -np.ones(n)
    g0[length] += n
    g0[m] += n
    g = g0 / (n ** 2 - 1 + (n % 2))
    w = np.fft.ihfft(v2 + g)
    w = w.real
    if n % 2 == 1:
        weights = np.concatenate([w, w[::-1]])
    else:
        weights = np.concatenate([w, w[len(w)-2::-1]])
    return points, weights


def build_integration_grid(quadm, gIntervals, phase):
    tg_n, ag_n = clenshaw_curtis(quadm)
    dstep = 2 * PI / float(gIntervals)
    dI = np.arange(-PI + 0.5 * dstep, PI, dstep)
    tg = (dI[:, None]+0.5*dstep*tg_n).ravel()
    ag = (dI[:, None]*0+ag_n).ravel()
    return dict(tg=tg, ag=ag, sintg=np.sin(tg), costg=np.cos(tg), sintgph=np.sin(tg + phase),
                costgph=np.cos(tg + phase), dstep=dstep)


def sp_sum(grid, Kx, Ky, ww1S, wS, wuS, gS, ddphiS, ddpsiS):
    sintg, costg, sintgph, costgph, tg, ag = (grid[k] for k in
                                              ('sintg', 'costg', 'sintgph', 'costgph', 'tg', 'ag'))
    sin2x = 2.*sintg*costg
    sin2xph = 2.*sintgph*costgph
    revgamma = 1. / gS
    revgamma2 = revgamma**2
    wwuS = wS/wuS
    Bsr = np.complex128(0)
    Bpr = np.complex128(0)
    dirx = ddphiS
    diry = ddpsiS
    dirz = 1. - 0.5*(ddphiS**2 + ddpsiS**2)
    taperC = 1
    alphaS = 0
    for i in range(len(tg)):
        ucos = ww1S*tg[i] + wwuS*revgamma *\
            (-Ky*ddphiS*sintg[i] + Kx*ddpsiS*sintgph[i] +
             0.125*revgamma*(Ky**2 * sin2x[i] + Kx**2 * sin2xph[i]))
        eucos = np.exp(1j*ucos)
        betax = taperC*Ky*revgamma*costg[i]
        betay = -Kx*revgamma*costgph[i]
        betaz = 1. - 0.5*(revgamma2 + betax*betax + betay*betay)
        betaPx = -Ky*(alphaS*costg[i] + taperC*sintg[i])
        betaPy = Kx*sintgph[i]
        betaPz = 0.5*revgamma *\
            (Ky**2 * taperC*(alphaS*costg[i]**2 + taperC*sin2x[i]) + Kx**2 * sin2xph[i])
        rkrel = 1./(1. - dirx*betax - diry*betay - dirz*betaz)
        eucos *= ag[i] * rkrel**2
        bnx = dirx - betax
        bny = diry - betay
        bnz = dirz - betaz
        dirDotBetaP = dirx*betaPx + diry*betaPy + dirz*betaPz
        dirDotDmB = dirx*bnx + diry*bny + dirz*bnz
        Bsr += eucos*(bnx*dirDotBetaP - betaPx*dirDotDmB)
        Bpr += eucos*(bny*dirDotBetaP - betaPy*dirDotDmB)
    return wuS*revgamma*Bsr, wuS*revgamma*Bpr


def sample_prefactors(L0, Np, Kx, Ky, gamma, w, ddtheta, ddpsi):
    """Per-sample wu, ww1, ab of _build_I_map_conv (synchr.py:2062-2071). gamma: array like w."""
    gamma2 = gamma**2
    wu = PI / L0 / gamma2 * np.ones_like(w) * (2*gamma2 - 1 - 0.5*Kx**2 - 0.5*Ky**2) / E2WC
    ww1 = w * ((1. + 0.5*Kx**2 + 0.5*Ky**2) + gamma2 * (ddtheta**2 + ddpsi**2)) / (2. * gamma2 * wu)
    ab = 1. / PI2 / wu * np.sin(PI * Np * ww1) / np.sin(PI * ww1)
    return wu, ww1, ab


def build_I_map(grid, L0, Np, Kx, Ky, eI, gamma, w, ddtheta, ddpsi, distE='eV'):
    """Returns (I, Es, Ep) of synchr.py:2104-2108 for per-sample gamma (array) -- the energy-spread draw
    (synchr.py:2053-2059) is the caller's business."""
    NRAYS = len(w)
    gamma = gamma * np.ones(NRAYS)
    wu, ww1, ab = sample_prefactors(L0, Np, Kx, Ky, gamma, w, ddtheta, ddpsi)
    Is_local, Ip_local = sp_sum(grid, Kx, Ky, ww1, w, wu, gamma, ddtheta, ddpsi)
    bwFact = 0.001 if distE == 'BW' else 1./w
    Amp2Flux = FINE_STR * bwFact * eI / SIE0
    integralField = np.abs(Is_local)**2 + np.abs(Ip_local)**2
    dstep = grid['dstep']
    return (Amp2Flux * ab**2 * 0.25 * dstep**2 * integralField,
            np.sqrt(Amp2Flux) * ab * Is_local * 0.5 * dstep,
            np.sqrt(Amp2Flux) * ab * Ip_local * 0.5 * dstep)
