"""Oracle: SourceFromField -- radiation of an electron in a tabulated magnetic field (TEST INFRASTRUCTURE ONLY).

Restates in numpy / scipy, in the reference's operation order (no filament beam):
  * SourceFromField._magnetic_field                 xrt/backends/raycing/sources/synchr.py:727-765
  * SourceFromField._build_integration_grid         synchr.py:988-1003
  * SourceFromField._build_trajectory_conv (RK4)    synchr.py:1049-1147
  * SourceFromField._sp_sum                         synchr.py:888-973 (far field and R0)
  * SourceFromField._build_I_map_custom_field_conv  synchr.py:1274-1347
  * read_custom_field / _fwhm (deviceLength)        synchr.py:695-710
"""
import numpy as np
from scipy.interpolate import interp1d

from .physconsts import SIE0, SIM0, C, E2WC, EMC, FINE_STR
from .undulator import clenshaw_curtis

SPL_KW = {'kind': 'cubic', 'bounds_error': False, 'fill_value': 'extrapolate'}


def fwhm(z, a):
    dz = z[1] - z[0]
    args = np.argwhere(a >= a.max()*0.5)
    return z[np.max(args)] - z[np.min(args)] + dz


def magnetic_field(data, grid):
    dataz = data[:, 0]
    ncol = data.shape[1]
    if ncol == 2:
        By = interp1d(dataz, data[:, 1], **SPL_KW)(grid)
        Bx = np.zeros_like(By)
        Bz = np.zeros_like(By)
    elif ncol == 3:
        Bx = interp1d(dataz, data[:, 1], **SPL_KW)(grid)
        By = interp1d(dataz, data[:, 2], **SPL_KW)(grid)
        Bz = np.zeros_like(By)
    else:
        Bx = interp1d(dataz, data[:, 1], **SPL_KW)(grid)
        By = interp1d(dataz, data[:, 2], **SPL_KW)(grid)
        Bz = interp1d(dataz, data[:, 3], **SPL_KW)(grid)
    return Bx, By, Bz


def grids(data):
    dataz = data[:, 0]
    lenmm = np.abs(dataz[-1] - dataz[0])
    wtGrid = np.linspace(dataz[0], dataz[-1], int(lenmm*10))
    BGrid = np.linspace(dataz[0], dataz[-1], 2*len(wtGrid)-1)
    return wtGrid, BGrid


def build_integration_grid(data, quadm, gIntervals):
    tg_n, ag_n = clenshaw_curtis(quadm)
    dataz = data[:, 0]
    dstep = (dataz[-1] - dataz[0]) / float(gIntervals)
    dI = np.arange(0.5 * dstep + dataz[0], dataz[-1], dstep)
    tg = (dI[:, None]+0.5*dstep*tg_n).ravel()
    ag = (dI[:, None]*0+ag_n).ravel()
    return tg, ag, dstep


def build_trajectory(wtGrid, tg, Bx, By, Bz):
    """_build_trajectory_conv, not filamentBeam: emcg = 1, betaz = -0.5 (bx^2 + by^2)."""
    emcg = 1.

    def f_beta(B, beta):
        return emcg*np.array((beta[1]*B[2]-B[1], B[0] - beta[0]*B[2]))

    def f_traj(beta):
        betaz = -0.5*(beta[0]**2 + beta[1]**2)
        return np.array((beta[0], beta[1], betaz))

    def next_beta_rk(iB, beta):
        k1beta = rkStep * f_beta([Bx[iB], By[iB], Bz[iB]], beta)
        k2beta = rkStep * f_beta([Bx[iB+1], By[iB+1], Bz[iB+1]], beta + 0.5*k1beta)
        k3beta = rkStep * f_beta([Bx[iB+1], By[iB+1], Bz[iB+1]], beta + 0.5*k2beta)
        k4beta = rkStep * f_beta([Bx[iB+2], By[iB+2], Bz[iB+2]], beta + k3beta)
        return beta + (k1beta + 2*k2beta + 2*k3beta + k4beta)/6.

    def next_traj_rk(iB, beta, traj):
        k1beta = rkStep * f_beta([Bx[iB], By[iB], Bz[iB]], beta)
        k1traj = rkStep * f_traj(beta)
        k2beta = rkStep * f_beta([Bx[iB+1], By[iB+1], Bz[iB+1]], beta + 0.5*k1beta)
        k2traj = rkStep * f_traj(beta + 0.5*k1beta)
        k3beta = rkStep * f_beta([Bx[iB+1], By[iB+1], Bz[iB+1]], beta + 0.5*k2beta)
        k3traj = rkStep * f_traj(beta + 0.5*k2beta)
        k4beta = rkStep * f_beta([Bx[iB+2], By[iB+2], Bz[iB+2]], beta + k3beta)
        k4traj = rkStep * f_traj(beta + k3beta)
        return (beta + (k1beta + 2*k2beta + 2*k3beta + k4beta)/6.,
                traj + (k1traj + 2*k2traj + 2*k3traj + k4traj)/6.)

    beta_next = np.zeros(2)
    beta0 = np.zeros(2)
    betam_int = 0
    for i in range(len(wtGrid)-1):
        rkStep = wtGrid[i+1] - wtGrid[i]
        beta_next = next_beta_rk(2*i, beta_next)
        beta0 += rkStep * beta_next
    beta0 /= -(wtGrid[-1] - wtGrid[0])
    beta_next = np.copy(beta0)
    traj_next = np.zeros(3)
    traj0 = np.zeros(3)
    for i in range(len(wtGrid)-1):
        rkStep = wtGrid[i+1] - wtGrid[i]
        beta_next, traj_next = next_traj_rk(2*i, beta_next, traj_next)
        traj0 += rkStep * traj_next
        betam_int += beta_next[0]**2 + beta_next[1]**2
    traj0 /= -(wtGrid[-1] - wtGrid[0])
    beta_next = np.copy(beta0)
    traj_next = np.copy(traj0)
    betam_int *= -0.5/(len(wtGrid)-1)
    betax = [beta0[0]]
    betay = [beta0[1]]
    trajx = [traj0[0]]
    trajy = [traj0[1]]
    trajz = [traj0[2]]
    for i in range(len(wtGrid)-1):
        rkStep = wtGrid[i+1] - wtGrid[i]
        beta_next, traj_next = next_traj_rk(2*i, beta_next, traj_next)
        betax.append(beta_next[0])
        betay.append(beta_next[1])
        trajx.append(traj_next[0])
        trajy.append(traj_next[1])
        trajz.append(traj_next[2])
    out = [interp1d(wtGrid, v, **SPL_KW)(tg) for v in (betax, betay, trajx, trajy, trajz)]
    return out[0], out[1], [betam_int], out[2], out[3], out[4]


def sp_sum(tg, ag, emcg, w, gamma, ddphi, ddpsi, Bx, By, Bz, betax, betay, betam, trajx, trajy, trajz, R0=None,
           with_scale=False):
    Bsr = np.complex128(0)
    Bpr = np.complex128(0)
    Ssr = 0.
    Spr = 0.
    gamma_ = gamma
    dirx = ddphi
    diry = ddpsi
    dirz = np.sqrt(1. - ddphi**2 - ddpsi**2)
    revgamma2 = 1./gamma_**2
    wc = w * E2WC / betam
    if R0 is not None:
        sinr0z, cosr0z = np.sin(wc*R0[2, :]), np.cos(wc*R0[2, :])
    for i in range(len(tg)):
        betax_ = emcg*betax[i]
        betay_ = emcg*betay[i]
        trajx_ = emcg*trajx[i]
        trajy_ = emcg*trajy[i]
        trajz_ = tg[i]*(1.-0.5*revgamma2) + EMC**2*revgamma2*trajz[i]
        if R0 is not None:
            rloc = np.array([trajx_, trajy_, trajz_])
            if len(rloc.shape) < 2:
                rloc = np.expand_dims(rloc, 1)
            dr = R0 - rloc
            dist = np.linalg.norm(dr, axis=0)
            rdrz = 1./dr[2, :]
            drs = (dr[0, :]**2+dr[1, :]**2)*rdrz
            LRS = 0.5*drs - 0.125*drs**2*rdrz + 0.0625*drs**3*rdrz**2
            sinzloc = np.sin(wc * (tg[i] - trajz_))
            coszloc = np.cos(wc * (tg[i] - trajz_))
            sindrs = np.sin(wc * LRS)
            cosdrs = np.cos(wc * LRS)
            eucosx = (-sinr0z*sinzloc*cosdrs - sinr0z*coszloc*sindrs -
                      cosr0z*sinzloc*sindrs + cosr0z*coszloc*cosdrs)
            eucosy = (-sinr0z*sinzloc*sindrs + sinr0z*coszloc*cosdrs +
                      cosr0z*sinzloc*cosdrs + cosr0z*coszloc*sindrs)
        else:
            phz = wc*(tg[i] - dirz*trajz_)
            phxy = wc*(dirx*trajx_ + diry*trajy_)
            sinphz, cosphz = np.sin(phz), np.cos(phz)
            sinphxy, cosphxy = np.sin(phxy), np.cos(phxy)
            eucosx = sinphz*cosphxy - cosphz*sinphxy
            eucosy = cosphz*cosphxy + sinphz*sinphxy
        eucos = eucosx + 1j*eucosy
        if R0 is not None:
            direction = dr/dist
            dirx = direction[0, :]
            diry = direction[1, :]
            dirz = direction[2, :]
        smTerm = revgamma2 + betax_**2 + betay_**2
        betaz = 1. - 0.5*smTerm - 0.125*smTerm**2 - 0.0625*smTerm**3
        betaPx = betay_*Bz[i] - betaz*By[i]
        betaPy = -betax_*Bz[i] + betaz*Bx[i]
        betaPz = betax_*By[i] - betay_*Bx[i]
        rkrel = 1./(1. - dirx*betax_ - diry*betay_ - dirz*betaz)
        eucos *= ag[i] * rkrel**2
        bnx = dirx - betax_
        bny = diry - betay_
        bnz = dirz - betaz
        dirDotBetaP = dirx*betaPx + diry*betaPy + dirz*betaPz
        dirDotDmB = dirx*bnx + diry*bny + dirz*bnz
        Bsr += eucos*(bnx*dirDotBetaP - betaPx*dirDotDmB)
        Bpr += eucos*(bny*dirDotBetaP - betaPy*dirDotDmB)
        if with_scale:  # test helper: sum of |terms|
            Ssr = Ssr + np.abs(eucos)*(np.abs(bnx*dirDotBetaP) + np.abs(betaPx*dirDotDmB))
            Spr = Spr + np.abs(eucos)*(np.abs(bny*dirDotBetaP) + np.abs(betaPy*dirDotDmB))
    if with_scale:
        return Bsr*emcg, Bpr*emcg, Ssr*emcg, Spr*emcg
    return Bsr*emcg, Bpr*emcg


class FieldSpec(object):
    """data: customFieldData (N x 2|3|4); gamma0, eI, quadm, gIntervals, distE, R0 (mm or None)."""

    def __init__(self, data, gamma0, eI, quadm, gIntervals, distE='eV', R0=None):
        self.data, self.gamma0, self.eI, self.distE, self.R0 = data, gamma0, eI, distE, R0
        self.wtGrid, self.BGrid = grids(data)
        self.tg, self.ag, self.dstep = build_integration_grid(data, quadm, gIntervals)
        Bx, By, Bz = magnetic_field(data, self.BGrid)
        self.traj = build_trajectory(self.wtGrid, self.tg, Bx, By, Bz)
        self.Bt = magnetic_field(data, self.tg)
        z = data[:, 0]
        self.deviceLength = fwhm(z, np.abs(data[:, 1:]).max(axis=1))

    def build_I_map(self, w, ddtheta, ddpsi, gamma, with_scale=False):
        """_build_I_map_custom_field_conv for a per-sample gamma array."""
        NRAYS = len(w)
        gamma = gamma * np.ones(NRAYS)
        R0 = self.R0 if self.R0 is not None else 0
        betax, betay, betazav, trajx, trajy, trajz = self.traj
        Bxt, Byt, Bzt = self.Bt
        betam = betazav[-1]
        ab = 0.5 / np.pi / (1. - 0.5/gamma**2 + betam*EMC**2/gamma**2)
        emcg = SIE0 / SIM0 / C / 10. / gamma
        if self.R0:
            R0v = np.array((np.tan(ddtheta), np.tan(ddpsi), np.ones_like(ddpsi)))
            R0v *= R0
        else:
            R0v = None
        res = sp_sum(self.tg, self.ag, emcg, w, gamma, ddtheta, ddpsi, Bxt, Byt, Bzt, betax, betay, betam, trajx,
                     trajy, trajz, R0v, with_scale)
        Is_local, Ip_local = res[0], res[1]
        bwFact = 0.001 if self.distE == 'BW' else 1./w
        Amp2Flux = FINE_STR * bwFact * self.eI / SIE0
        integralField = np.abs(Is_local)**2 + np.abs(Ip_local)**2
        out = (Amp2Flux * 0.25 * self.dstep**2 * ab**2 * integralField,
               np.sqrt(Amp2Flux) * Is_local * 0.5 * self.dstep * ab,
               np.sqrt(Amp2Flux) * Ip_local * 0.5 * self.dstep * ab)
        if with_scale:
            amp = np.abs(np.sqrt(Amp2Flux) * 0.5 * self.dstep * ab)
            return out + (amp * res[2], amp * res[3])
        return out
