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


class Wave(object):
    """The fields of a prepared wave that shine(wave=...) reads and writes (waves.prepare_wave, waves.py:505-584)."""

    def __init__(self, xDiffr, yDiffr, zDiffr, area, areaNormal=None):
        n = len(xDiffr)
        self.xDiffr, self.yDiffr, self.zDiffr = np.array(xDiffr), np.array(yDiffr), np.array(zDiffr)
        self.area = area
        if areaNormal is not None:
            self.areaNormal = areaNormal
        b = ob.Beam(n, withAmplitudes=True)
        self.__dict__.update({k: v for k, v in b.__dict__.items() if k not in self.__dict__})
        self.rDiffr = (self.xDiffr**2 + self.yDiffr**2 + self.zDiffr**2)**0.5
        self.a[:] = self.xDiffr / self.rDiffr
        self.b[:] = self.yDiffr / self.rDiffr
        self.c[:] = self.zDiffr / self.rDiffr
        self.Jss[:] = 0
        self.Jpp[:] = 0


def shine_wave(src, wave, filamentBeam=False, toGlobal=True, record=None, R0=None, accuBeam=None):
    """IntegratedSource.shine(wave=wave) with uniformRayDensity (sybase.py:1513-1812): the receiving points of *wave*
    (in source-local coordinates xDiffr, yDiffr, zDiffr) are illuminated; *wave* gets E, directions, J, Es, Ep (with
    the propagation phase), the returned beam sits at the source.  Consumes np.random in the reference's order and
    records the deviates."""
    from .physconsts import CHBAR
    assert src.uniformRayDensity
    mcRays = len(wave.a)
    rec = {}
    dgamma = None
    seeded0, seededI0 = np.int64(0), 0.
    if filamentBeam and accuBeam is not None:  # the same electron as in an earlier shine (:1553-1561)
        rsE = accuBeam.E[0]
        rX, rZ = accuBeam.filamentDX, accuBeam.filamentDZ
        dtheta, dpsi = accuBeam.filamentDtheta, accuBeam.filamentDpsi
        dgamma = accuBeam.filamentDgamma
        seeded0, seededI0 = accuBeam.seeded, accuBeam.seededI
    elif filamentBeam:  # :1541-1552
        rec['uE'] = np.random.random_sample()
        rsE = rec['uE'] * float(src.E_max - src.E_min) + src.E_min
        rec['gX'], rec['gZ'] = np.random.standard_normal(), None
        rX = src.dx * rec['gX']
        rec['gZ'] = np.random.standard_normal()
        rZ = src.dz * rec['gZ']
        rec['gT'] = np.random.standard_normal()
        dtheta = src.dxprime * rec['gT']
        rec['gP'] = np.random.standard_normal()
        dpsi = src.dzprime * rec['gP']
        if getattr(src, 'eEspread', 0) > 0:
            rec['gG'] = np.random.standard_normal()
            dgamma = src.gamma * src.eEspread * rec['gG']
    seeded = seeded0 + mcRays
    if filamentBeam:
        rE = rsE * np.ones(mcRays)
    else:
        rec['uE'] = np.random.rand(mcRays)
        rE = rec['uE'] * float(src.E_max - src.E_min) + src.E_min
    src.xzE = (src.E_max - src.E_min)  # :1589
    if filamentBeam:
        shiftX, shiftZ = rX, rZ
    else:
        rec['gX'] = np.random.standard_normal(mcRays) if src.dx > 0 else None
        shiftX = 0. + src.dx * rec['gX'] if src.dx > 0 else 0
        rec['gZ'] = np.random.standard_normal(mcRays) if src.dz > 0 else None
        shiftZ = 0. + src.dz * rec['gZ'] if src.dz > 0 else 0
    x = wave.xDiffr + shiftX
    y = wave.yDiffr
    z = wave.zDiffr + shiftZ
    rDiffr = np.sqrt((x**2 + y**2 + z**2))
    rTheta = x / rDiffr
    rPsi = z / rDiffr
    if filamentBeam:
        rTheta += dtheta
        rPsi += dpsi
    else:
        if src.dxprime > 0:
            rec['gT'] = np.random.standard_normal(mcRays)
            rTheta += 0. + src.dxprime * rec['gT']
        if src.dzprime > 0:
            rec['gP'] = np.random.standard_normal(mcRays)
            rPsi += 0. + src.dzprime * rec['gP']
    gamma = src.gamma
    if getattr(src, 'eEspread', 0) > 0:  # synchr.py:2053-2059
        if dgamma is not None:
            gamma = gamma + dgamma
        else:
            sz = 1 if filamentBeam else mcRays
            rec['g5'] = np.random.normal(size=sz)
            gamma = gamma + gamma * src.eEspread * rec['g5']
    rec['gamma'] = gamma * np.ones(mcRays)
    Intensity, mJs, mJp = ou.build_I_map(src.grid, src.L0, src.Np, src.Kx, src.Ky, src.eI, gamma * np.ones(mcRays),
                                         rE, rTheta, rPsi, src.distE)
    seededI = seededI0 + mcRays * src.xzE
    sourceWeight = src.xzE
    tmp_max = np.max(Intensity)
    if tmp_max > src.Imax:
        src.Imax = tmp_max
        src.fluxConst = src.Imax * src.xzE
    bot = wave
    bot.state[:] = 1
    bot.E[:] = rE
    if filamentBeam:
        dxR, dzR = rX, rZ
    else:
        SX, SZ = src.get_SIGMA(bot.E)
        rec['g1'] = np.random.standard_normal(mcRays)
        dxR = 0. + SX * rec['g1']
        rec['g2'] = np.random.standard_normal(mcRays)
        dzR = 0. + SZ * rec['g2']
    wave.rDiffr = np.sqrt(((wave.xDiffr - dxR)**2 + wave.yDiffr**2 + (wave.zDiffr - dzR)**2))
    wave.path[:] = 0
    wave.a[:] = (wave.xDiffr - dxR) / wave.rDiffr
    wave.b[:] = wave.yDiffr / wave.rDiffr
    wave.c[:] = (wave.zDiffr - dzR) / wave.rDiffr
    area = wave.areaNormal if hasattr(wave, 'areaNormal') else wave.area
    norm = area**0.5 / wave.rDiffr
    mJs = mJs * norm
    mJp = mJp * norm
    mJs2 = (mJs * np.conj(mJs)).real
    mJp2 = (mJp * np.conj(mJp)).real
    sSP = 1.
    bot.Jsp[:] = np.where(sSP, mJs * np.conj(mJp) / sSP, 0)
    bot.Jss[:] = np.where(sSP, mJs2 / sSP, 0)
    bot.Jpp[:] = np.where(sSP, mJp2 / sSP, 0)
    bot.Es[:] = mJs
    bot.Ep[:] = mJp
    bo = bot
    length = len(bo.a)
    bo.accepted = length * src.fluxConst
    bo.acceptedE = bo.E.sum() * src.fluxConst * SIE0
    bo.seeded = seeded
    bo.seededI = seededI
    bo.sourceWeight = sourceWeight / len(wave.a)
    if filamentBeam:
        bo.filamentDtheta, bo.filamentDpsi, bo.filamentDX, bo.filamentDZ = dtheta, dpsi, rX, rZ
        bo.filamentDgamma = dgamma
    norm = (bo.a**2 + bo.b**2 + bo.c**2)**0.5
    bo.a /= norm
    bo.b /= norm
    bo.c /= norm
    if src.pitch or src.yaw:
        ob.rotate_beam(bo, pitch=src.pitch, yaw=src.yaw)
    bor = ob.Beam(copyFrom=bo)
    for k in ('accepted', 'acceptedE', 'seeded', 'seededI', 'sourceWeight'):
        setattr(bor, k, getattr(bo, k))
    bor.x[:] = dxR
    bor.y[:] = 0.
    bor.z[:] = dzR
    if R0 is None:
        bor.path[:] = 0.
        mPh = np.exp(1e7j * wave.E/CHBAR * wave.rDiffr)
        wave.Es *= mPh
        wave.Ep *= mPh
    if toGlobal:
        ob.virgin_local_to_global(src.azimuth, bor, src.center)
    if record is not None:
        record.update(rec)
    return bor
