"""Oracle: BendingMagnet / Wiggler sources (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Restates in numpy + scipy.special.kv, in the reference's operation order:
  * BendingMagnet.build_I_map   xrt/backends/raycing/sources/synchr.py:185-226 (also Wiggler: isMPW branch :204-206)
  * BendingMagnet.shine         synchr.py:229-509 (no filament beam, no fixedEnergy / accuBeam)
`shine()` consumes the legacy global `np.random` stream in exactly the reference's order, so that after
`np.random.seed(s)` it reproduces the reference's beam bit for bit; `record` receives the deviates it drew so
that the CUDA path can be fed the same numbers."""
import numpy as np
from scipy import special

from . import beam as ob
from .physconsts import PI, PI2, SIE0, SIM0, E2W, FINE_STR, CHeVcm, SQ3


class BMSpec(object):
    """Parameter bag: gamma, B, Np, isMPW, K, L0, X0, ro, eI, eEspread, window, sizes, placement."""

    def __init__(self, **kw):
        self.center = [0., 0., 0.]
        self.azimuth = (0., 1.)
        self.pitch = self.yaw = 0.
        self.uniformRayDensity = False
        self.distE = 'eV'
        self.eEspread = 0.
        self.__dict__.update(kw)
        self.Imax = 0.
        self.xzE = (self.E_max - self.E_min) * (self.Theta_max - self.Theta_min) * (self.Psi_max - self.Psi_min)
        self.fluxConst = self.Imax * self.xzE

    def build_I_map(self, dde, ddtheta, ddpsi, record=None):
        gamma = self.gamma
        if self.eEspread > 0:
            if np.array(dde).shape:
                if dde.shape[0] > 1:
                    g6 = np.random.standard_normal(dde.shape)  # normal(0, s, shape) == 0 + s * standard_normal
                    gamma = gamma + (0. + gamma*self.eEspread * g6)
                    if record is not None:
                        record['g6'] = g6
            gamma2 = gamma**2
        else:
            gamma2 = self.gamma**2
        return build_I_map(self, gamma, gamma2, dde, ddtheta, ddpsi)


def build_I_map(src, gamma, gamma2, dde, ddtheta, ddpsi):
    """synchr.py:199-226 for given gamma (scalar or per-sample array)."""
    with np.errstate(invalid='ignore', divide='ignore', over='ignore'):
        w_cr = 1.5 * gamma2 * src.B * SIE0 / SIM0
        if src.isMPW:
            thetaArg = ddtheta * gamma / src.K
            w_cr *= np.sqrt(np.clip(1. - thetaArg**2, 0., None))
        w_cr = np.where(np.isfinite(w_cr), w_cr, 0.)
        gammapsi = gamma * ddpsi
        gamma2psi2p1 = gammapsi**2 + 1
        eta = 0.5 * dde * E2W / w_cr * gamma2psi2p1**1.5
        ampSP = -0.5j * SQ3 / PI * gamma * dde * E2W / w_cr * gamma2psi2p1
        ampS = ampSP * special.kv(2./3., eta)
        ampP = 1j * gammapsi * ampSP * special.kv(1./3., eta) / np.sqrt(gamma2psi2p1)
        ampS = np.where(np.isfinite(ampS), ampS, 0.)
        ampP = np.where(np.isfinite(ampP), ampP, 0.)
        bwFact = 0.001 if src.distE == 'BW' else 1./dde
        Amp2Flux = FINE_STR * bwFact * src.eI / SIE0 * 2 * src.Np
        return (Amp2Flux * (np.abs(ampS)**2 + np.abs(ampP)**2),
                np.sqrt(Amp2Flux) * ampS,
                np.sqrt(Amp2Flux) * ampP)


def shine(src, nrays, toGlobal=True, withAmplitudes=True, record=None):
    if src.uniformRayDensity:
        withAmplitudes = True
    bo = None
    length = 0
    seeded = np.int64(0)
    seededI = 0.
    mcRays = np.int64(nrays * 1.2) if not src.uniformRayDensity else nrays
    passes = []
    rep_condition = True
    while rep_condition:
        rnd_r = np.random.rand(mcRays, 4)
        seeded += mcRays
        rE = rnd_r[:, 0] * float(src.E_max - src.E_min) + src.E_min
        rTheta = (rnd_r[:, 1]) * (src.Theta_max - src.Theta_min) + src.Theta_min
        rPsi = rnd_r[:, 2] * (src.Psi_max - src.Psi_min) + src.Psi_min
        extra = {}
        Intensity, mJss, mJpp = src.build_I_map(rE, rTheta, rPsi, extra)
        if src.uniformRayDensity:
            seededI += nrays * src.xzE
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
        else:
            I_pass = np.where(src.Imax * rnd_r[:, 3] < Intensity)[0]
            npassed = len(I_pass)
        if npassed == 0:
            continue
        bot = ob.Beam(npassed, withAmplitudes=withAmplitudes)
        bot.state[:] = 1
        bot.E[:] = rE[I_pass]
        Theta0 = rTheta[I_pass]
        Psi0 = rPsi[I_pass]
        rec = dict(E=rE, theta=rTheta, psi=rPsi, u=rnd_r[:, 3].copy(), I_pass=I_pass, Imax=src.Imax, **extra)
        if src.dxprime > 0:
            rec['g1'] = np.random.standard_normal(npassed)
            dtheta = 0. + src.dxprime * rec['g1']
        else:
            dtheta = 0
        if not src.isMPW:
            rec['g2'] = np.random.standard_normal(npassed)
            dtheta += 0. + 1/src.gamma * rec['g2']
        if src.dzprime > 0:
            rec['g3'] = np.random.standard_normal(npassed)
            dpsi = 0. + src.dzprime * rec['g3']
        else:
            dpsi = 0
        bot.a[:] = np.tan(Theta0 + dtheta)
        bot.c[:] = np.tan(Psi0 + dpsi)
        intensS = (mJss[I_pass] * np.conj(mJss[I_pass])).real
        intensP = (mJpp[I_pass] * np.conj(mJpp[I_pass])).real
        if src.uniformRayDensity:
            sSP = 1.
        else:
            sSP = intensS + intensP
        if src.isMPW:
            sigma_r2 = 2 * (CHeVcm/bot.E*10*src.L0*src.Np) / PI2**2
            sourceSIGMAx = np.sqrt(src.dx**2 + sigma_r2)
            sourceSIGMAz = np.sqrt(src.dz**2 + sigma_r2)
            thetaArg = np.clip(Theta0*src.gamma/src.K, -1., 1.)
            rec['k'] = np.random.randint(-int(src.Np), int(src.Np), npassed)
            bot.y[:] = ((np.arccos(thetaArg) / PI) + rec['k'] - 0.5) * 0.5 * src.L0
            rec['g4'] = np.random.standard_normal(npassed)
            bot.x[:] = src.X0 * np.sin(PI2 * bot.y / src.L0) + (0. + sourceSIGMAx * rec['g4'])
            rec['g5'] = np.random.standard_normal(npassed)
            bot.z[:] = 0. + sourceSIGMAz * rec['g5']
            bot.Jsp[:] = np.zeros(npassed)
        else:
            if src.dz > 0:
                rec['g5'] = np.random.standard_normal(npassed)
                bot.z[:] = 0. + src.dz * rec['g5']
            if src.dx > 0:
                rec['g4'] = np.random.standard_normal(npassed)
                R1 = src.ro*1e3 + src.dx * rec['g4']
            else:
                R1 = src.ro * 1e3
            bot.x[:] = -R1 * np.cos(Theta0) + src.ro*1000.
            bot.y[:] = R1 * np.sin(Theta0)
            with np.errstate(divide='ignore', invalid='ignore'):
                bot.Jsp[:] = np.array(np.where(sSP, mJss[I_pass] * np.conj(mJpp[I_pass]) / sSP, sSP), dtype=complex)
        with np.errstate(divide='ignore', invalid='ignore'):
            bot.Jss[:] = np.where(sSP, intensS / sSP, sSP)
            bot.Jpp[:] = np.where(sSP, intensP / sSP, sSP)
        if withAmplitudes:
            bot.Es[:] = mJss[I_pass]
            bot.Ep[:] = mJpp[I_pass]
        passes.append(rec)
        if bo is None:
            bo = bot
        else:
            for k, v in list(bo.__dict__.items()):
                if isinstance(v, np.ndarray):
                    setattr(bo, k, np.concatenate((v, getattr(bot, k))))
        length = len(bo.a)
        rep_condition = length < nrays
        if src.uniformRayDensity:
            rep_condition = False
    if length >= nrays:
        bo.accepted = length * src.fluxConst
        bo.acceptedE = bo.E.sum() * src.fluxConst * SIE0
        bo.seeded = seeded
        bo.seededI = seededI
        bo.sourceWeight = sourceWeight / nrays
    if length > nrays:
        for k, v in list(bo.__dict__.items()):
            if isinstance(v, np.ndarray):
                setattr(bo, k, v[:nrays])
    norm = np.sqrt(bo.a**2 + 1.0 + bo.c**2)
    bo.a /= norm
    bo.b /= norm
    bo.c /= norm
    if src.pitch or src.yaw:
        ob.rotate_beam(bo, pitch=src.pitch, yaw=src.yaw)
    if toGlobal:
        ob.virgin_local_to_global(src.azimuth, bo, src.center)
    if record is not None:
        record.extend(passes)
    return bo
