"""Oracle: undulator far-field integral (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Restates, in numpy and in the reference's operation order:
  * IntegratedSource._clenshaw_curtis        xrt/backends/raycing/sources/sybase.py:1106-1139
  * Undulator._build_integration_grid        xrt/backends/raycing/sources/synchr.py:1787-1801
  * Undulator._sp_sum                        synchr.py:1929-2033: far field (the `else` branch :1997-2003),
                                             taper (:1958-1971) and near field R0 (:1972-1996) + :2005-2031
  * Undulator._build_I_map_conv              synchr.py:2048-2108
  * SourceBase.intensities_on_mesh           sybase.py:740-928 (resultKind='Stokes')
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


def sp_sum(grid, Kx, Ky, ww1S, wS, wuS, gS, ddphiS, ddpsiS, R0=None, taperVal=None, Np=1, with_scale=False):
    """R0: (3, n) array in units of L0/2pi (synchr.py:2072-2077) or None; taperVal: Undulator._taperVal or
    None; Np: number of periods (the explicit period loop runs only with R0 or taper, synchr.py:1947).
    with_scale (test helper, not in the reference): also return sum_j |term_j| for Bs and Bp, the honest
    scale of an oscillatory quadrature."""
    sintg, costg, sintgph, costgph, tg, ag = (grid[k] for k in
                                              ('sintg', 'costg', 'sintgph', 'costgph', 'tg', 'ag'))
    taperC = 1
    alphaS = 0
    sin2x = 2.*sintg*costg
    sin2xph = 2.*sintgph*costgph
    revgamma = 1. / gS
    revgamma2 = revgamma**2
    betam = 1. - (1. + 0.5*Kx**2 + 0.5*Ky**2)*0.5*revgamma2
    wwuS = wS/wuS
    Bsr = np.complex128(0)
    Bpr = np.complex128(0)
    Ssr = 0.
    Spr = 0.
    dirx = ddphiS
    diry = ddpsiS
    dirz = 1. - 0.5*(ddphiS**2 + ddpsiS**2)
    Nmx = int(Np) if (R0 is not None or taperVal is not None) else 1
    if R0 is not None:
        sinr0z = np.sin(R0[-1])
        cosr0z = np.cos(R0[-1])
    for Nperiod in range(Nmx):
        for i in range(len(tg)):
            if taperVal is not None:
                zloc = -(Nmx-1)*np.pi + Nperiod*PI2 + tg[i]
                alphaS = taperVal/E2WC
                taperC = 1. - alphaS*zloc/wuS
                ucos = ww1S*zloc +\
                    wwuS*revgamma *\
                    (-Ky*dirx*(sintg[i] + alphaS/wuS *
                               (1 - costg[i] -
                                zloc*sintg[i])) +
                     Kx*diry*sintg[i] + 0.125*revgamma *
                     (Kx**2 * sin2xph[i] + Ky**2 * (sin2x[i] -
                      2*alphaS/wuS*(zloc**2 + costg[i]**2 +
                                    zloc*sin2x[i]))))
                eucos = np.exp(1j*ucos)
            elif R0 is not None:
                zterm = 0.5*(Ky**2*sin2x[i] + Kx**2*sin2xph[i])*revgamma
                zloc = -(Nmx-1)*np.pi + Nperiod*PI2 + tg[i]
                rloc = np.array([Ky*sintg[i]*revgamma,
                                 Kx*sintgph[i]*revgamma,
                                 betam*zloc-0.25*zterm*revgamma])
                dr = R0 - rloc
                dist = np.linalg.norm(dr, axis=0)
                drs = 0.5*(dr[0, :]**2+dr[1, :]**2) / dr[2, :]
                sinzloc = np.sin(wwuS * zloc*(1.-betam))
                coszloc = np.cos(wwuS * zloc*(1.-betam))
                sindrs = np.sin(wwuS * (drs + 0.25*zterm*revgamma))
                cosdrs = np.cos(wwuS * (drs + 0.25*zterm*revgamma))
                eucosx = (-sinr0z*sinzloc*cosdrs - sinr0z*coszloc*sindrs -
                          cosr0z*sinzloc*sindrs + cosr0z*coszloc*cosdrs)
                eucosy = (-sinr0z*sinzloc*sindrs + sinr0z*coszloc*cosdrs +
                          cosr0z*sinzloc*cosdrs + cosr0z*coszloc*sindrs)
                eucos = eucosx + 1j*eucosy
                direction = dr/dist
                dirx = direction[0, :]
                diry = direction[1, :]
                dirz = direction[2, :]
            else:
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
            if with_scale:
                Ssr = Ssr + np.abs(eucos)*(np.abs(bnx*dirDotBetaP) + np.abs(betaPx*dirDotDmB))
                Spr = Spr + np.abs(eucos)*(np.abs(bny*dirDotBetaP) + np.abs(betaPy*dirDotDmB))
    if with_scale:
        return wuS*revgamma*Bsr, wuS*revgamma*Bpr, wuS*revgamma*Ssr, wuS*revgamma*Spr
    return wuS*revgamma*Bsr, wuS*revgamma*Bpr


def sample_prefactors(L0, Np, Kx, Ky, gamma, w, ddtheta, ddpsi):
    """Per-sample wu, ww1, ab of _build_I_map_conv (synchr.py:2062-2071). gamma: array like w."""
    gamma2 = gamma**2
    wu = PI / L0 / gamma2 * np.ones_like(w) * (2*gamma2 - 1 - 0.5*Kx**2 - 0.5*Ky**2) / E2WC
    ww1 = w * ((1. + 0.5*Kx**2 + 0.5*Ky**2) + gamma2 * (ddtheta**2 + ddpsi**2)) / (2. * gamma2 * wu)
    ab = 1. / PI2 / wu * np.sin(PI * Np * ww1) / np.sin(PI * ww1)
    return wu, ww1, ab


def build_I_map(grid, L0, Np, Kx, Ky, eI, gamma, w, ddtheta, ddpsi, distE='eV', R0=None, taperVal=None):
    """Returns (I, Es, Ep) of synchr.py:2104-2108 for per-sample gamma (array) -- the energy-spread draw
    (synchr.py:2053-2059) is the caller's business.  R0 [mm] / taperVal switch to the explicit period sum with
    ab = 1/(2 pi wu) (synchr.py:2067-2079)."""
    NRAYS = len(w)
    gamma = gamma * np.ones(NRAYS)
    wu, ww1, ab = sample_prefactors(L0, Np, Kx, Ky, gamma, w, ddtheta, ddpsi)
    if (taperVal is not None) or (R0 is not None):
        ab = 1. / PI2 / wu
    if R0:
        R0v = np.array((np.tan(ddtheta), np.tan(ddpsi), np.ones_like(ddpsi)))
        R0v *= R0*np.pi*2/L0
    else:
        R0v = None
    Is_local, Ip_local = sp_sum(grid, Kx, Ky, ww1, w, wu, gamma, ddtheta, ddpsi, R0v, taperVal, Np)
    bwFact = 0.001 if distE == 'BW' else 1./w
    Amp2Flux = FINE_STR * bwFact * eI / SIE0
    integralField = np.abs(Is_local)**2 + np.abs(Ip_local)**2
    dstep = grid['dstep']
    return (Amp2Flux * ab**2 * 0.25 * dstep**2 * integralField,
            np.sqrt(Amp2Flux) * ab * Is_local * 0.5 * dstep,
            np.sqrt(Amp2Flux) * ab * Ip_local * 0.5 * dstep)


def intensities_on_mesh(grid, L0, Np, Kx, Ky, eI, gamma0, eEspread, dxprime, dzprime, energy, theta, psi,
                        harmonic=None, eSpreadSigmas=3.5, eSpreadNSamples=36, mode='constant', distE='eV'):
    """SourceBase.intensities_on_mesh (sybase.py:740-928), Stokes result, for the far-field undulator."""
    tomesh = [energy, theta, psi]
    if harmonic is not None:
        tomesh.append(harmonic)
        iharmonic = len(tomesh)-1
    else:
        iharmonic = None
    if eEspread > 0:
        spr = np.linspace(-eSpreadSigmas, eSpreadSigmas, eSpreadNSamples)
        dgamma = gamma0 * spr * eEspread
        wspr = np.exp(-0.5 * spr**2)
        wspr /= wspr.sum()
        tomesh.append(dgamma)
        ispread = len(tomesh)-1
    else:
        ispread = None
    mesh = np.meshgrid(*tomesh, indexing='ij')
    xE, xTheta, xPsi = mesh[0].ravel(), mesh[1].ravel(), mesh[2].ravel()
    sh = [len(energy), len(theta), len(psi)]
    if iharmonic:
        xH = mesh[iharmonic].ravel()
        sh.append(len(harmonic))
    else:
        xH = None
    gamma = gamma0
    if ispread:
        gamma = gamma + mesh[ispread].ravel()  # synchr.py:2053-2056
        sh.append(len(dgamma))
    NRAYS = len(xE)
    gamma = gamma * np.ones(NRAYS)
    wu, ww1, ab = sample_prefactors(L0, Np, Kx, Ky, gamma, xE, xTheta, xPsi)
    Is_local, Ip_local = sp_sum(grid, Kx, Ky, ww1, xE, wu, gamma, xTheta, xPsi)
    bwFact = 0.001 if distE == 'BW' else 1./xE
    Amp2Flux = FINE_STR * bwFact * eI / SIE0
    if xH is not None:  # synchr.py:2094-2098
        Is_local[ww1 > xH+0.5] = 0
        Ip_local[ww1 > xH+0.5] = 0
        Is_local[ww1 < xH-0.5] = 0
        Ip_local[ww1 < xH-0.5] = 0
    dstep = grid['dstep']
    Es = (np.sqrt(Amp2Flux) * ab * Is_local * 0.5 * dstep).reshape(sh)
    Ep = (np.sqrt(Amp2Flux) * ab * Ip_local * 0.5 * dstep).reshape(sh)
    Is = (Es*np.conj(Es)).real.astype(float)
    Ip = (Ep*np.conj(Ep)).real.astype(float)
    Isp = Es*np.conj(Ep).astype(complex)
    if ispread:
        if iharmonic:
            ws = wspr[np.newaxis, np.newaxis, np.newaxis, np.newaxis, :]
        else:
            ws = wspr[np.newaxis, np.newaxis, np.newaxis, :]
        Is = (Is * ws).sum(axis=ispread)
        Ip = (Ip * ws).sum(axis=ispread)
        Isp = (Isp * ws).sum(axis=ispread)
    s0 = Is + Ip
    s1 = Is - Ip
    s2 = 2. * np.real(Isp)
    s3 = -2. * np.imag(Isp)
    ss = [s0, s1, s2, s3]
    if (dxprime > 0 or dzprime > 0) and len(theta) > 1 and len(psi) > 1:
        from scipy.ndimage import gaussian_filter
        Sx = dxprime / (theta[1] - theta[0])
        Sz = dzprime / (psi[1] - psi[0])
        for ie, ee in enumerate(energy):
            if harmonic is None:
                for arr in ss:
                    arr[ie, :, :] = gaussian_filter(arr[ie, :, :], [Sx, Sz], mode=mode)
            else:
                for ih, hh in enumerate(harmonic):
                    for arr in ss:
                        arr[ie, :, :, ih] = gaussian_filter(arr[ie, :, :, ih], [Sx, Sz], mode=mode)
    with np.errstate(divide='ignore', invalid='ignore'):
        return [s0, np.where(s0, s1/s0, s0), np.where(s0, s2/s0, s0), np.where(s0, s3/s0, s0)]
