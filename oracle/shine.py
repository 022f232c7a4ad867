"""Oracle: IntegratedSource.shine for the far-field undulator (TEST INFRASTRUCTURE ONLY).

Restates xrt/backends/raycing/sources/sybase.py:1469-1812 (no wave, no filament beam, eEspread = 0) in numpy.
`shine()` consumes the legacy global `np.random` stream in exactly the reference's order, so that after
`np.random.seed(s)` it reproduces the reference's beam bit for bit; `record` receives the deviates it drew so
that the CUDA path can be fed the same numbers."""
import numpy as np

from . import beam as ob, undulator as ou
from .physconsts import PI2, CHeVcm, SIE0, PI


class UndulatorSpec(object):
    def __init__(self, **kw):
        self.center = [0., 0., 0.]
        self.azimuth = (0., 1.)
        self.pitch = self.yaw = 0.
        self.uniformRayDensity = False
        self.distE = 'eV'
        self.__dict__.update(kw)
        self.Imax = 0.
        self.xzE = (self.E_max - self.E_min) * (self.Theta_max - self.Theta_min) * (self.Psi_max - self.Psi_min)
        self.fluxConst = self.Imax * self.xzE
        self.grid = ou.build_integration_grid(self.quadm, self.gIntervals, self.phase)

    def build_I_map(self, w, th, ps, record=None):
        gamma = self.gamma
        if getattr(self, 'eEspread', 0) > 0:  # synchr.py:2053-2059
            g5 = np.random.normal(size=len(w))
            gamma = gamma + gamma * self.eEspread * g5
            if record is not None:
                record['g5'] = g5
        return ou.build_I_map(self.grid, self.L0, self.Np, self.Kx, self.Ky, self.eI, gamma * np.ones(len(w)),
                              w, th, ps, self.distE)

    def get_SIGMA(self, E):  # synchr.py:2200-2249 (onlyOddHarmonics=False as shine passes it), sybase.py:664-674
        sigma_r2 = 2 * CHeVcm/E*10 * self.L0*self.Np / PI2**2
        if getattr(self, 'eEspread', 0) > 0:
            from scipy import special
            harmonic = np.floor_divide(E, self.E1)
            eEspread_norm = PI2 * harmonic * self.Np * self.eEspread
            x = eEspread_norm/4.
            ret = np.ones_like(x, dtype=float)
            xarr = np.array(x)
            y = 2**0.5 * xarr[xarr > 1e-6]
            y2 = y**2
            ret[x > 1e-6] = y2 / (np.exp(-y2) + PI**0.5*y*special.erf(y) - 1)
            sigma_r2 = sigma_r2 * ret**(2/3.)
        return (self.dx**2 + sigma_r2)**0.5, (self.dz**2 + sigma_r2)**0.5


def shine(src, nrays, toGlobal=True, withAmplitudes=True, record=None):
    mcRays = nrays
    if src.uniformRayDensity:
        withAmplitudes = True
    bo = None
    length = 0
    seeded = np.int64(0)
    seededI = 0.
    passes = []
    while True:
        seeded += mcRays
        rndg = np.random.rand(mcRays)
        rE = rndg * float(src.E_max - src.E_min) + src.E_min
        rndg = np.random.rand(mcRays)
        rTheta = rndg * (src.Theta_max - src.Theta_min) + src.Theta_min
        rndg = np.random.rand(mcRays)
        rPsi = rndg * (src.Psi_max - src.Psi_min) + src.Psi_min
        extra = {}
        Intensity, mJs, mJp = src.build_I_map(rE, rTheta, rPsi, extra)
        if src.uniformRayDensity:
            seededI += mcRays * src.xzE
            sourceWeight = src.xzE
        else:
            seededI += Intensity.sum() * src.xzE
            sourceWeight = seededI / seeded
        tmp_max = np.max(Intensity)
        if tmp_max > src.Imax:
            src.Imax = tmp_max
            src.fluxConst = src.Imax * src.xzE
        if src.uniformRayDensity:
            I_pass = slice(None)
            npassed = mcRays
            u = None
        else:
            u = np.random.rand(mcRays)
            I_pass = np.where(src.Imax * u < Intensity)[0]
            npassed = len(I_pass)
        if npassed == 0:
            continue
        bot = ob.Beam(npassed, withAmplitudes=withAmplitudes)
        bot.state[:] = 1
        bot.E[:] = rE[I_pass]
        SX, SZ = src.get_SIGMA(bot.E)
        g1 = np.random.standard_normal(npassed)  # np.random.normal(0, SX, n) == 0 + SX * standard_normal
        dxR = 0. + SX * g1
        g2 = np.random.standard_normal(npassed)
        dzR = 0. + SZ * g2
        bot.x[:] = dxR
        bot.z[:] = dzR
        bot.a[:] = rTheta[I_pass]
        bot.c[:] = rPsi[I_pass]
        g3 = g4 = None
        if src.dxprime > 0:
            g3 = np.random.standard_normal(npassed)
            bot.a[:] += 0. + src.dxprime * g3
        if src.dzprime > 0:
            g4 = np.random.standard_normal(npassed)
            bot.c[:] += 0. + src.dzprime * g4
        passes.append(dict(E=rE, theta=rTheta, psi=rPsi, u=u, I_pass=I_pass, g1=g1, g2=g2, g3=g3, g4=g4,
                           Imax=src.Imax, **extra))
        mJs = mJs[I_pass]
        mJp = mJp[I_pass]
        mJs2 = (mJs * np.conj(mJs)).real
        mJp2 = (mJp * np.conj(mJp)).real
        sSP = 1. if src.uniformRayDensity else mJs2 + mJp2
        with np.errstate(divide='ignore', invalid='ignore'):
            bot.Jsp[:] = np.where(sSP, mJs * np.conj(mJp) / sSP, 0)
            bot.Jss[:] = np.where(sSP, mJs2 / sSP, 0)
            bot.Jpp[:] = np.where(sSP, mJp2 / sSP, 0)
        if withAmplitudes:
            if src.uniformRayDensity:
                bot.Es[:] = mJs
                bot.Ep[:] = mJp
            else:
                bot.Es[:] = mJs / mJs2**0.5
                bot.Ep[:] = mJp / mJp2**0.5
        if bo is None:
            bo = bot
        else:
            for k, v in list(bo.__dict__.items()):
                if isinstance(v, np.ndarray):
                    setattr(bo, k, np.concatenate((v, getattr(bot, k))))
        length = len(bo.a)
        bo.accepted = length * src.fluxConst
        bo.acceptedE = bo.E.sum() * src.fluxConst * SIE0
        bo.seeded = seeded
        bo.seededI = seededI
        bo.sourceWeight = sourceWeight / nrays
        if src.uniformRayDensity or length >= nrays:
            break
    if length > nrays:
        for k, v in list(bo.__dict__.items()):
            if isinstance(v, np.ndarray):
                setattr(bo, k, v[:nrays])
    norm = (bo.a**2 + bo.b**2 + bo.c**2)**0.5
    bo.a /= norm
    bo.b /= norm
    bo.c /= norm
    if src.pitch or src.yaw:
        ob.rotate_beam(bo, pitch=src.pitch, yaw=src.yaw)
    bor = ob.Beam(copyFrom=bo)
    if toGlobal:
        ob.virgin_local_to_global(src.azimuth, bor, src.center)
    if record is not None:
        record.extend(passes)
    return bor
