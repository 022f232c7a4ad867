# -*- coding: utf-8 -*-
"""`Undulator` with xrt's constructor and `shine()` / `build_I_map()` signatures, computing on the GPU.

Mirrors xrt/backends/raycing/sources/sybase.py (SourceBase :29-547, IntegratedSource :933-1812) and
sources/synchr.py (Undulator :1349-2260) for planar/helical undulators in the far field, with a linear gap
taper (`taper=(dgap, gap)`) or in the near field (`R0`): same keywords and units (eE GeV, eI A, eSigma um, eEpsilon nm rad, xPrimeMax mrad, period mm), same derived
quantities (gamma, dx/dz/dxprime/dzprime, Theta/Psi/E windows, Clenshaw-Curtis grid), same flux bookkeeping
on the returned Beam (accepted, acceptedE, seeded, seededI, sourceWeight).  Out of scope:
filamentBeam, custom field, wave=... sampling.

Per-ray work: xrt_b200/csrc/undulator.cu (field integral), source.cu (sampling, Beam assembly).
"""
import ctypes

import numpy as np
import torch

from . import _lib, kernels
from .beams import Beam
from .cuda_backend import XRT_CUDA
from .physconsts import PI, PI2, EV2ERG, M0, C, E2WC, FINE_STR, SIE0, K2B, EMC, CHBAR, CHeVcm

F64 = torch.float64
C128 = torch.complex128


class xrt_source_params(ctypes.Structure):
    """struct xrt_source_params of include/xrt_b200.h"""
    _fields_ = [('center', ctypes.c_double * 3), ('sinAz', ctypes.c_double), ('cosAz', ctypes.c_double),
                ('to_global', ctypes.c_int), ('uniform', ctypes.c_int),
                ('dx', ctypes.c_double), ('dz', ctypes.c_double), ('dxprime', ctypes.c_double),
                ('dzprime', ctypes.c_double), ('L0', ctypes.c_double), ('Np', ctypes.c_double),
                ('has_pitch', ctypes.c_int), ('has_yaw', ctypes.c_int),
                ('cosPitch', ctypes.c_double), ('sinPitch', ctypes.c_double),
                ('cosYaw', ctypes.c_double), ('sinYaw', ctypes.c_double),
                ('eEspread', ctypes.c_double), ('E1', ctypes.c_double), ('gamma0', ctypes.c_double)]


def clenshaw_curtis(n):
    """Nodes and weights on [-1, 1] (IntegratedSource._clenshaw_curtis, sybase.py:1106-1139, after quadpy)."""
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
    w = np.fft.ihfft(v2 + g).real
    if n % 2 == 1:
        weights = np.concatenate([w, w[::-1]])
    else:
        weights = np.concatenate([w, w[len(w)-2::-1]])
    return points, weights


class Undulator(object):
    def __init__(self, bl=None, name='und', center=(0, 0, 0), nrays=100000, eE=6.0, eI=0.1, eEspread=0.,
                 eSigmaX=None, eSigmaZ=None, eEpsilonX=1., eEpsilonZ=0.01, betaX=9., betaZ=2., eMin=9000,
                 eMax=9100, distE='eV', xPrimeMax=0.01, zPrimeMax=0.01, uniformRayDensity=False,
                 filamentBeam=False, pitch=0, yaw=0, period=50, n=50, K=1, Kx=0, Ky=0, phaseDeg=0, gp=1e-6,
                 gIntervals=2, gNodes=None, xPrimeMaxAutoReduce=True, zPrimeMaxAutoReduce=True,
                 targetOpenCL='CUDA', precisionOpenCL='float64', R0=None, taper=None, seed=1234, **kwargs):
        if targetOpenCL is None:
            raise _lib.XrtB200Error('xrt_b200.Undulator always computes on the GPU (no numpy path)')
        # a filament beam (one electron per shine) is built for the wave-mode entry shine(wave=...); the ray-mode
        # filament loop (nrepmax repeats, sybase.py:1742-1744) stays outside the hot path
        self.filamentBeam = bool(filamentBeam)
        self.bl = bl
        self.name = name
        if bl is not None and self not in bl.sources:
            bl.sources.append(self)
        self.center = list(center)
        self.nrays = int(nrays)
        self.pitch, self.yaw = pitch, yaw
        self.distE = distE
        self.uniformRayDensity = uniformRayDensity
        self.eE = float(eE)
        self.gamma = self.eE * 1e9 * EV2ERG / (M0 * C**2)  # sybase.py:155
        self.gamma2 = self.gamma**2
        self.eEspread = eEspread
        self.eI = float(eI)
        # emittance [nm rad] -> [mm rad]; sizes [um] -> [mm]; beta [m] -> [mm] (sybase.py:218-347)
        epsX = eEpsilonX * 1e-6 if eEpsilonX is not None else None
        epsZ = eEpsilonZ * 1e-6 if eEpsilonZ is not None else None

        def size(sigma, eps, beta):
            if sigma is not None:
                d = sigma * 1e-3
            elif eps is not None and beta is not None:
                d = np.sqrt(eps * beta * 1e3)
            else:
                d = 0
            dprime = (eps / d if d > 0 else 0) if eps is not None else 0
            return d, dprime
        self.dx, self.dxprime = size(eSigmaX, epsX, betaX)
        self.dz, self.dzprime = size(eSigmaZ, epsZ, betaZ)
        self.eMin, self.eMax = float(eMin), float(eMax)
        self.L0 = float(period)
        self.Np = float(n)
        self.R0 = R0  # near field: distance to the screen [mm] (sybase.py:84-149)
        self._set_taper(taper)
        if self.R0 is not None:  # required for convergence (synchr.py:1433-1435)
            xPrimeMaxAutoReduce = zPrimeMaxAutoReduce = True
        self.phase = np.radians(phaseDeg)
        if Kx == 0 and Ky == 0:  # synchr.py:1414-1431
            self.Kx, self.Ky = 0., float(K)
        else:
            self.Kx, self.Ky = float(Kx), float(Ky)
        self.B0x, self.B0y = K2B * self.Kx / self.L0, K2B * self.Ky / self.L0
        self._xPrimeMax = self._prime(xPrimeMax)
        self._zPrimeMax = self._prime(zPrimeMax)
        if xPrimeMaxAutoReduce:  # sybase.py:372-386
            K0 = self.Ky if abs(self.Ky) > 0 else 2.
            self._xPrimeMax = [np.sign(v) * min(abs(v), K0 / self.gamma) for v in self._xPrimeMax]
        if zPrimeMaxAutoReduce:
            K0 = self.Kx if abs(self.Kx) > 0 else 2.
            self._zPrimeMax = [np.sign(v) * min(abs(v), K0 / self.gamma) for v in self._zPrimeMax]
        self.gp = gp
        self.gIntervals = int(gIntervals)
        self.quadm = int(gNodes) if gNodes is not None else None
        self.maxIntegrationNodes = int(6e5)
        self.seed = int(seed)
        self._draws = 0
        self.ucl = XRT_CUDA('undulator.cl', targetOpenCL, precisionOpenCL)
        self.cl_ctx = self.ucl.cl_ctx
        self.cl_precisionF, self.cl_precisionC = self.ucl.cl_precisionF, self.ucl.cl_precisionC
        self.report_E1()
        self.needReset = True

    def _set_taper(self, taper):
        """Undulator.taper setter (synchr.py:1566-1594): a number is dB/dx/B-like `_taperVal` itself, a pair is
        (dgap, gap) [mm] -> _taperVal = dgap / Np / L0 / gap."""
        if taper is None:
            self._taperVal = None
            self.taper = None
        elif np.ndim(taper) == 0:
            self._taperVal = float(taper)
            self.taper = self._taperVal
        else:
            t = np.asarray(taper, dtype=float).ravel()
            if len(t) == 1:
                dgap, gap = 0., t[0]
            elif len(t) == 2:
                dgap, gap = t
            else:
                raise ValueError('Cannot set taper from {0}'.format(taper))
            self.taper = (dgap, gap)
            self.gap = gap
            self._taperVal = None if dgap == 0 else dgap / self.Np / self.L0 / gap

    @property
    def _ext_variant(self):
        """(variant, scalar0) of the explicit-period kernels, the dispatch of synchr.py:2136-2170; None: far field."""
        if self._taperVal is not None:
            return 1, self._taperVal
        if self.R0 is not None:
            return 2, self.R0*np.pi*2/self.L0
        return None

    @staticmethod
    def _prime(v):
        if isinstance(v, (tuple, list)):
            return [min(v) * 1e-3, max(v) * 1e-3]
        return [-abs(v) * 1e-3, abs(v) * 1e-3]

    def report_E1(self):  # synchr.py:1661-1677
        wu = PI / self.L0 / self.gamma2 * (2*self.gamma2 - 1. - 0.5*self.Kx**2 - 0.5*self.Ky**2) / E2WC
        self.E1 = 2*wu*self.gamma2 / (1 + 0.5*self.Kx**2 + 0.5*self.Ky**2)

    # ------------------------------------------------------------------ reset / integration grid
    def reset(self):
        self.needReset = False
        # limits corrected for divergence (sybase.py:479-515)
        self.Theta_min = float(self._xPrimeMax[0] - self.dxprime)
        self.Theta_max = float(self._xPrimeMax[1] + self.dxprime)
        self.Psi_min = float(self._zPrimeMax[0] - self.dzprime)
        self.Psi_max = float(self._zPrimeMax[1] + self.dzprime)
        self.E_min = float(min(self.eMin, self.eMax))
        self.E_max = float(max(self.eMin, self.eMax))
        if self.quadm is None:
            self._find_convergence_mixed()
        self._build_integration_grid()
        self.Imax = 0.
        if self.filamentBeam and not hasattr(self, 'dimExy'):
            # sybase.py:529-541: Imax probed with nrays uniform samples (no energy spread), 20 % head room
            dev = torch.device('cuda', torch.cuda.current_device())
            gen = torch.Generator(device=dev)
            gen.manual_seed(self.seed * 7919 + 17)
            u = torch.rand((4, self.nrays), dtype=torch.float64, device=dev, generator=gen)
            DistI = self.build_I_map_device(u[0] * (self.E_max - self.E_min) + self.E_min,
                                            u[1] * (self.Theta_max - self.Theta_min) + self.Theta_min,
                                            u[2] * (self.Psi_max - self.Psi_min) + self.Psi_min)[0]
            self.Imax = float(DistI.max()) * 1.2
            self.nrepmax = np.floor(self.nrays / max(int((self.Imax * u[3] < DistI).sum()), 1))
        self.xzE = (self.E_max - self.E_min) * (self.Theta_max - self.Theta_min) * (self.Psi_max - self.Psi_min)
        self.fluxConst = self.Imax * self.xzE

    def get_SIGMA_device(self, E):
        """(SIGMAx, SIGMAz) per sample: get_SIGMA with onlyOddHarmonics=False as shine passes it (synchr.py:2200-2249,
        sybase.py:664-674), CUDA tensors."""
        sigma_r2 = 2 * CHeVcm/E*10 * self.L0*self.Np / PI2**2
        if self.eEspread > 0:
            harmonic = torch.floor(E / self.E1)
            x = (PI2 * harmonic * self.Np * self.eEspread) / 4.
            y = 2**0.5 * x
            y2 = y**2
            qa = y2 / (torch.exp(-y2) + np.pi**0.5 * y * torch.special.erf(y) - 1)
            ret = torch.where(x > 1e-6, qa, torch.ones_like(x))
            sigma_r2 = sigma_r2 * ret**(2/3.)
        return (self.dx**2 + sigma_r2)**0.5, (self.dz**2 + sigma_r2)**0.5

    def shine_wave(self, wave, toGlobal=True, samples=None, accuBeam=None):
        """IntegratedSource.shine(wave=wave) (sybase.py:1513-1527, :1588-1612, :1651-1703, :1786-1796): the receiving
        points of a prepared wave (Screen / aperture / OE .prepare_wave with this source as the previous element) are
        illuminated; *wave* receives E, directions, coherency matrix and Es / Ep with the propagation phase, the
        returned Beam sits at the source.  The field of every sample is the CUDA `build_I_map`; the O(M) rest is
        device elementwise work.  *samples*: optional dict of deviates (parity tests feed the oracle's): filament
        beam: scalars uE, gX, gZ, gT, gP[, gG]; full emittance: CUDA tensors uE, gX, gZ, gT, gP, g1, g2[, g5]."""
        if not hasattr(wave, 'rDiffr'):
            raise ValueError("If you want to use a `wave`, run a `prepare_wave` before shine!")
        if self.needReset:
            self.reset()
        self.uniformRayDensity = True  # sybase.py:1523
        n = len(wave)
        dev = wave.device
        s = samples
        seeded0, seededI0 = 0, 0.
        if accuBeam is not None:  # the same electron as in an earlier shine (sybase.py:1553-1561)
            if not self.filamentBeam:
                raise NotImplementedError('accuBeam goes with filamentBeam=True')
            E0 = float(accuBeam.E[0])
            s = dict(uE=None, gX=None, gZ=None, gT=None, gP=None, gG=None)
            seeded0, seededI0 = accuBeam.seeded, accuBeam.seededI
        if s is None:
            gen = torch.Generator(device=dev)
            gen.manual_seed(self.seed * 104729 + self._draws)
            self._draws += 1
            if self.filamentBeam:
                u = torch.rand(1, dtype=torch.float64, device=dev, generator=gen)
                g = torch.randn(5, dtype=torch.float64, device=dev, generator=gen).tolist()
                s = dict(uE=float(u), gX=g[0], gZ=g[1], gT=g[2], gP=g[3], gG=g[4])
            else:
                g = torch.randn((7, n), dtype=torch.float64, device=dev, generator=gen)
                s = dict(uE=torch.rand(n, dtype=torch.float64, device=dev, generator=gen), gX=g[0], gZ=g[1], gT=g[2],
                         gP=g[3], g1=g[4], g2=g[5], g5=g[6])
        dgamma = None
        if self.filamentBeam and accuBeam is not None:
            rX, rZ = accuBeam.filamentDX, accuBeam.filamentDZ
            dtheta, dpsi = accuBeam.filamentDtheta, accuBeam.filamentDpsi
            dgamma = accuBeam.filamentDgamma
            rE = torch.full((n,), E0, dtype=torch.float64, device=dev)
            shiftX, shiftZ = rX, rZ
        elif self.filamentBeam:  # one electron: energy, position, direction [and energy deviation] (sybase.py:1541-1552)
            rX, rZ = self.dx * s['gX'], self.dz * s['gZ']
            dtheta, dpsi = self.dxprime * s['gT'], self.dzprime * s['gP']
            if self.eEspread > 0:
                dgamma = self.gamma * self.eEspread * s['gG']
            rE = torch.full((n,), s['uE'] * float(self.E_max - self.E_min) + self.E_min, dtype=torch.float64,
                            device=dev)
            shiftX, shiftZ = rX, rZ
        else:
            rE = s['uE'] * float(self.E_max - self.E_min) + self.E_min
            shiftX = 0. + self.dx * s['gX'] if self.dx > 0 else 0
            shiftZ = 0. + self.dz * s['gZ'] if self.dz > 0 else 0
        self.xzE = (self.E_max - self.E_min)  # sybase.py:1589
        x = wave.xDiffr + shiftX
        y = wave.yDiffr
        z = wave.zDiffr + shiftZ
        rDiffr = torch.sqrt(x**2 + y**2 + z**2)
        rTheta = x / rDiffr
        rPsi = z / rDiffr
        if self.filamentBeam:
            rTheta = rTheta + dtheta
            rPsi = rPsi + dpsi
        else:
            if self.dxprime > 0:
                rTheta = rTheta + (0. + self.dxprime * s['gT'])
            if self.dzprime > 0:
                rPsi = rPsi + (0. + self.dzprime * s['gP'])
        gam = None
        if self.eEspread > 0:  # synchr.py:2053-2059
            if dgamma is not None:
                gam = torch.full((n,), self.gamma + dgamma, dtype=torch.float64, device=dev)
            else:
                gam = self.gamma + (self.gamma * self.eEspread) * s['g5']
        Intensity, mJs, mJp = self.build_I_map_device(rE.contiguous(), rTheta.contiguous(), rPsi.contiguous(),
                                                      gamma=gam)
        seeded = seeded0 + n
        seededI = seededI0 + n * self.xzE
        tmp_max = float(Intensity.max())
        if tmp_max > self.Imax:
            self.Imax = tmp_max
            self.fluxConst = self.Imax * self.xzE
        wave.state = torch.ones_like(wave.state)
        wave.E = rE
        if self.filamentBeam:
            dxR, dzR = rX, rZ
        else:
            SX, SZ = self.get_SIGMA_device(rE)
            wave.sourceSIGMAx, wave.sourceSIGMAz = SX, SZ
            dxR = 0. + SX * s['g1']
            dzR = 0. + SZ * s['g2']
        wave.rDiffr = torch.sqrt((wave.xDiffr - dxR)**2 + wave.yDiffr**2 + (wave.zDiffr - dzR)**2)
        wave.path = torch.zeros_like(wave.path)
        a = (wave.xDiffr - dxR) / wave.rDiffr
        b = wave.yDiffr / wave.rDiffr
        c = (wave.zDiffr - dzR) / wave.rDiffr
        area = wave.areaNormal if hasattr(wave, 'areaNormal') else wave.area
        norm = area**0.5 / wave.rDiffr
        mJs = mJs * norm
        mJp = mJp * norm
        wave.Jsp = mJs * torch.conj(mJp)
        wave.Jss = (mJs * torch.conj(mJs)).real.contiguous()
        wave.Jpp = (mJp * torch.conj(mJp)).real.contiguous()
        wave.accepted = n * self.fluxConst
        wave.acceptedE = float(rE.sum()) * self.fluxConst * SIE0
        wave.seeded = seeded
        wave.seededI = seededI
        wave.sourceWeight = self.xzE / n
        if self.filamentBeam:
            wave.filamentDtheta, wave.filamentDpsi, wave.filamentDX, wave.filamentDZ = dtheta, dpsi, rX, rZ
            wave.filamentDgamma = dgamma
        norm = (a**2 + b**2 + c**2)**0.5
        wave.a, wave.b, wave.c = a / norm, b / norm, c / norm
        if self.pitch or self.yaw:
            raise NotImplementedError('shine(wave=...) with a pitched / yawed source is outside the hot path')
        wave.Es, wave.Ep = mJs, mJp
        bor = Beam(copyFrom=wave)
        for k in ('accepted', 'acceptedE', 'seeded', 'seededI', 'sourceWeight'):
            setattr(bor, k, getattr(wave, k))
        bor.x = torch.zeros_like(wave.x) + dxR
        bor.y = torch.zeros_like(wave.y)
        bor.z = torch.zeros_like(wave.z) + dzR
        if self.R0 is None:  # sybase.py:1791-1796
            mPh = torch.exp(1e7j * rE / CHBAR * wave.rDiffr)
            wave.Es = wave.Es * mPh
            wave.Ep = wave.Ep * mPh
        bor.parentId = getattr(self, 'uuid', self.name)
        if toGlobal:  # virgin_local_to_global (beamline.py:262-287)
            a0 = self.bl.sinAzimuth if self.bl is not None else 0.
            b0 = self.bl.cosAzimuth if self.bl is not None else 1.
            if a0 != 0:
                bor.a, bor.b = b0*bor.a - (-a0)*bor.b, (-a0)*bor.a + b0*bor.b
                bor.x, bor.y = b0*bor.x - (-a0)*bor.y, (-a0)*bor.x + b0*bor.y
            bor.x = bor.x + self.center[0]
            bor.y = bor.y + self.center[1]
            bor.z = bor.z + self.center[2]
        return bor

    def _build_integration_grid(self):
        """Undulator._build_integration_grid (synchr.py:1787-1801)."""
        self.tg_n, self.ag_n = clenshaw_curtis(self.quadm)
        dstep = 2 * PI / float(self.gIntervals)
        dI = np.arange(-PI + 0.5 * dstep, PI, dstep)
        self.tg = (dI[:, None]+0.5*dstep*self.tg_n).ravel()
        self.ag = (dI[:, None]*0+self.ag_n).ravel()
        self.sintg = np.sin(self.tg)
        self.costg = np.cos(self.tg)
        self.sintgph = np.sin(self.tg + self.phase)
        self.costgph = np.cos(self.tg + self.phase)
        self.dstep = dstep
        self._dev_tables = None
        # a planar undulator on a mirror-symmetric grid lets the kernel visit node pairs (csrc/undulator.cu)
        self.gridSymmetric = kernels.grid_is_symmetric(self.tg, self.ag, self.sintg, self.costg)

    def _tables(self, device):
        if self._dev_tables is None or self._dev_tables[0].device != device:
            self._dev_tables = [torch.from_numpy(np.ascontiguousarray(a)).to(device) for a in
                                (self.tg, self.ag, self.sintg, self.costg, self.sintgph, self.costgph)]
        return self._dev_tables

    def _get_mad(self):
        """IntegratedSource._get_mad (sybase.py:1245-1285) at (E_max, Theta_max, Psi_max)."""
        tmp_quadm = self.quadm
        k = self.quadm - 2
        pltout, dIout = [], []
        Iold = None
        for m in range(5):
            k += 1
            self.quadm = k
            self._build_integration_grid()
            Inew = self._convergence_value()
            if m > 0:
                pltout.append(Inew)
                dIout.append(np.abs(Inew-Iold)/Inew)
            Iold = Inew
        med = np.median(np.abs(np.array(pltout)))
        mad = np.median(np.abs(np.abs(np.array(pltout)) - med))
        self.quadm = tmp_quadm
        return mad, np.median(dIout)

    def _convergence_value(self):
        """build_I_map with convergenceSearchFlag (synchr.py:2101-2103): |sqrt(|Is|^2+|Ip|^2) * dstep/2| for the
        single corner sample, eEspread = 0; evaluated by the same CUDA kernel."""
        w = np.array([self.E_max, self.E_max])  # two identical samples: the kernel path is the array path
        th = np.array([self.Theta_max, self.Theta_max])
        ps = np.array([self.Psi_max, self.Psi_max])
        gamma = self.gamma * np.ones(2)
        wu, ww1 = self._wu_ww1(gamma, w, th, ps)
        Is, Ip = self._integral_host(gamma, wu, w, ww1, th, ps)
        field = np.abs(Is[0])**2 + np.abs(Ip[0])**2
        return np.abs(np.sqrt(field) * 0.5 * self.dstep)

    def _find_convergence_mixed(self):
        """IntegratedSource._find_convergence_mixed (sybase.py:1190-1243)."""
        m = 3
        while m < 10000:
            m += 1
            self.quadm = int(2**m)
            mad, dimad = self._get_mad()
            if (dimad < self.gp) or (mad < self.gp):
                break
            if self.quadm > self.maxIntegrationNodes:
                break
        ph2start = int(2**(m-1))
        ph2end = self.quadm
        jmax = int(np.log2((ph2end-ph2start) / (4*5)))
        for j in range(jmax):
            self.quadm = int(0.5*(ph2end+ph2start))
            mad, dimad = self._get_mad()
            if (dimad < self.gp) or (mad < self.gp):
                ph2end = self.quadm
            else:
                ph2start = self.quadm
        self.quadm = ph2end

    # ------------------------------------------------------------------ build_I_map (host arrays, reference API)
    def _wu_ww1(self, gamma, w, ddtheta, ddpsi):
        gamma2 = gamma**2
        wu = PI / self.L0 / gamma2 * (2*gamma2 - 1 - 0.5*self.Kx**2 - 0.5*self.Ky**2) / E2WC
        ww1 = w * ((1. + 0.5 * self.Kx**2 + 0.5 * self.Ky**2) +
                   gamma2 * (ddtheta * ddtheta + ddpsi * ddpsi)) / (2. * gamma2 * wu)
        return wu, ww1

    def _integral_host(self, gamma, wu, w, ww1, ddtheta, ddpsi):
        """The dispatch of Undulator._build_I_map_CL (synchr.py:2134-2174) through XRT_CUDA.run_parallel."""
        F = self.cl_precisionF
        NRAYS = len(w)
        scalarArgs = [F(0.)]
        if self.R0 is not None:
            scalarArgs = [F(self.R0*np.pi*2/self.L0)]
        elif self._taperVal:
            scalarArgs = [F(self._taperVal)]
        scalarArgs.extend([F(self.Kx), F(self.Ky), np.int32(len(self.tg))])
        if (self._taperVal is not None) or (self.R0 is not None):
            scalarArgs.extend([np.int32(self.Np)])
        slicedROArgs = [F(gamma), F(wu), F(w), F(ww1), F(ddtheta), F(ddpsi)]
        nonSlicedROArgs = [F(self.tg), F(self.ag), F(self.sintg), F(self.costg), F(self.sintgph), F(self.costgph)]
        slicedRWArgs = [np.zeros(NRAYS, dtype=self.cl_precisionC), np.zeros(NRAYS, dtype=self.cl_precisionC)]
        if self._taperVal is not None:
            clKernel = 'undulator_taper'
        elif self.R0 is not None:
            clKernel = 'undulator_nf'
        else:
            clKernel = 'undulator'
        return self.ucl.run_parallel(clKernel, scalarArgs, slicedROArgs, nonSlicedROArgs, slicedRWArgs,
                                     dimension=NRAYS)

    def build_I_map(self, w, ddtheta, ddpsi, harmonic=None, dg=None):
        """(I, Es, Ep) on host arrays like Undulator.build_I_map (synchr.py:2035-2193)."""
        if self.needReset:
            self.reset()
        w = np.atleast_1d(np.asarray(w, dtype=float))
        ddtheta = np.asarray(ddtheta, dtype=float) * np.ones_like(w)
        ddpsi = np.asarray(ddpsi, dtype=float) * np.ones_like(w)
        NRAYS = len(w)
        gamma = self.gamma
        if self.eEspread > 0:
            if dg is not None:
                gamma = gamma + dg
            else:
                gamma = gamma + gamma * self.eEspread * np.random.normal(size=NRAYS)
        gamma = gamma * np.ones(NRAYS)
        wu, ww1 = self._wu_ww1(gamma, w, ddtheta, ddpsi)
        if (self._taperVal is not None) or (self.R0 is not None):  # synchr.py:2129-2132
            ab = 1. / PI2 / wu
        else:
            ab = 1. / PI2 / wu * np.sin(PI * self.Np * ww1) / np.sin(PI * ww1)
        Is_local, Ip_local = self._integral_host(gamma, wu, w, ww1, ddtheta, ddpsi)
        bwFact = 0.001 if self.distE == 'BW' else 1./w
        Amp2Flux = FINE_STR * bwFact * self.eI / SIE0
        if harmonic is not None:
            Is_local[ww1 > harmonic+0.5] = 0
            Ip_local[ww1 > harmonic+0.5] = 0
            Is_local[ww1 < harmonic-0.5] = 0
            Ip_local[ww1 < harmonic-0.5] = 0
        integralField = np.abs(Is_local)**2 + np.abs(Ip_local)**2
        return (Amp2Flux * ab**2 * 0.25 * self.dstep**2 * integralField,
                np.sqrt(Amp2Flux) * Is_local * ab * 0.5 * self.dstep,
                np.sqrt(Amp2Flux) * Ip_local * ab * 0.5 * self.dstep)

    # ------------------------------------------------------------------ device-resident building blocks
    def build_I_map_device(self, w, theta, psi, gamma=None, out=None):
        """(I, Es, Ep) as CUDA tensors from CUDA tensors (fused _build_I_map_conv)."""
        if self.needReset:
            self.reset()
        ext = self._ext_variant
        if ext is not None:
            return kernels.undulator_ext_imap(ext[0], ext[1], self.Kx, self.Ky, self.L0, self.Np, self.gamma,
                                              FINE_STR * self.eI / SIE0, self.dstep,
                                              1 if self.distE == 'BW' else 0, w, theta, psi,
                                              self._tables(w.device), gamma=gamma, out=out)
        return kernels.undulator_imap(self.Kx, self.Ky, self.L0, self.Np, self.gamma,
                                      FINE_STR * self.eI / SIE0, self.dstep, 1 if self.distE == 'BW' else 0,
                                      w, theta, psi, self._tables(w.device), gamma=gamma, out=out,
                                      symmetric=self.gridSymmetric)

    def _source_params(self, toGlobal):
        p = xrt_source_params()
        for i in range(3):
            p.center[i] = float(self.center[i])
        p.sinAz = float(self.bl.sinAzimuth) if self.bl is not None else 0.
        p.cosAz = float(self.bl.cosAzimuth) if self.bl is not None else 1.
        p.to_global = int(bool(toGlobal))
        p.uniform = int(bool(self.uniformRayDensity))
        p.dx, p.dz, p.dxprime, p.dzprime = float(self.dx), float(self.dz), float(self.dxprime), float(self.dzprime)
        p.L0, p.Np = self.L0, self.Np
        if self.pitch:
            p.has_pitch, p.cosPitch, p.sinPitch = 1, float(np.cos(self.pitch)), float(np.sin(self.pitch))
        if self.yaw:
            p.has_yaw, p.cosYaw, p.sinYaw = 1, float(np.cos(self.yaw)), float(np.sin(self.yaw))
        p.eEspread, p.E1, p.gamma0 = float(self.eEspread), float(self.E1), float(self.gamma)
        return p

    def draw_samples(self, n, device=None, index_offset=0, withNormals=True, withU=True):
        """Philox-sampled (E, theta, psi), four standard normal deviates and a uniform per ray (CUDA tensors)."""
        if self.needReset:
            self.reset()
        lib = _lib.load()
        dev = torch.device('cuda', torch.cuda.current_device()) if device is None else device
        names = ['E', 'theta', 'psi'] + (['n1', 'n2', 'n3', 'n4'] if withNormals else []) + (['u'] if withU else []) \
            + (['n5'] if self.eEspread > 0 else [])
        buf = torch.empty((len(names), n), dtype=F64, device=dev)
        s = {k: buf[i] for i, k in enumerate(names)}
        window = (ctypes.c_double * 6)(self.E_min, self.E_max, self.Theta_min, self.Theta_max, self.Psi_min,
                                       self.Psi_max)
        ptr = lambda k: s[k].data_ptr() if k in s else None  # noqa: E731
        # every call advances the counter space so that successive passes never reuse deviates
        offset = int(index_offset) + self._draws
        self._draws += 1 << 40
        rc = lib.xrt_source_sample_f64(n, self.seed, offset, window, ptr('E'), ptr('theta'), ptr('psi'), ptr('n1'),
                                       ptr('n2'), ptr('n3'), ptr('n4'), ptr('u'), ptr('n5'),
                                       ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, 'xrt_source_sample_f64')
        return s

    def fill_beam(self, samples, mJs, mJp, out, idx=None, out_offset=0, n=None, toGlobal=True):
        lib = _lib.load()
        p = self._source_params(toGlobal)
        n = (len(idx) if idx is not None else samples['E'].numel()) if n is None else n
        cout = out.c_struct()
        rc = lib.xrt_source_fill_f64(
            n, ctypes.byref(p), idx.data_ptr() if idx is not None else None, samples['E'].data_ptr(),
            samples['theta'].data_ptr(), samples['psi'].data_ptr(), mJs.data_ptr(), mJp.data_ptr(),
            samples['n1'].data_ptr(), samples['n2'].data_ptr(), samples['n3'].data_ptr(), samples['n4'].data_ptr(),
            ctypes.byref(cout), out_offset, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, 'xrt_source_fill_f64')

    def shine_uniform_fused(self, out, index_offset=0, toGlobal=True):
        """uniformRayDensity source in ONE kernel (sample + field integral + Beam assembly fused; bit-identical to
        draw_samples -> build_I_map_device -> fill_beam).  Returns the device slot holding max(I) of this call."""
        lib = _lib.load()
        if self.needReset:
            self.reset()
        dev = out.device
        slot = torch.zeros(1, dtype=torch.int64, device=dev)
        tabs = self._tables(dev)
        p = self._source_params(toGlobal)
        cout = out.c_struct()
        window = (ctypes.c_double * 6)(self.E_min, self.E_max, self.Theta_min, self.Theta_max, self.Psi_min,
                                       self.Psi_max)
        offset = int(index_offset) + self._draws
        self._draws += 1 << 40
        rc = lib.xrt_shine_uniform_f64(
            len(out), self.seed, offset, window, len(self.tg), self.Kx, self.Ky, self.L0, self.Np, self.gamma,
            FINE_STR * self.eI / SIE0, self.dstep, 1 if self.distE == 'BW' else 0, int(self.gridSymmetric),
            *[t.data_ptr() for t in tabs], ctypes.byref(p), ctypes.byref(cout), slot.data_ptr(),
            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, 'xrt_shine_uniform_f64')
        return slot

    # ------------------------------------------------------------------ shine
    def shine(self, toGlobal=True, withAmplitudes=True, fixedEnergy=False, wave=None, accuBeam=None,
              samples=None, index_offset=0, allreduce_max=None, bookkeeping=True, fused=True):
        """Returns the source Beam (device resident) like IntegratedSource.shine (sybase.py:1469-1812).

        *samples* (optional): dict of CUDA tensors E, theta, psi, n1..n4[, u] to use instead of drawing them
        (parity tests feed the deviates the numpy oracle used).  *allreduce_max*: callable applied to the
        per-pass max intensity when rays are sharded over ranks (Imax is beam-wide, sybase.py:1627-1631)."""
        if fixedEnergy or (accuBeam is not None and wave is None):
            raise NotImplementedError('fixedEnergy / ray-mode accuBeam sampling is outside the hot path (SURVEY 8f3)')
        if wave is not None:
            return self.shine_wave(wave, toGlobal=toGlobal, samples=samples, accuBeam=accuBeam)
        if self.filamentBeam:
            raise NotImplementedError('a filament beam is built for shine(wave=...); its ray-mode repeats '
                                      '(sybase.py:1742-1744) are outside the hot path')
        if self.needReset:
            self.reset()
        if self.uniformRayDensity:
            withAmplitudes = True
        mcRays = self.nrays
        dev = torch.device('cuda', torch.cuda.current_device())
        bo = Beam(mcRays, withAmplitudes=withAmplitudes, device=dev, uninitialized=True)
        length = 0
        seeded = 0
        seededI = 0.
        sumE = None
        sourceWeight = 0.
        given = samples is not None
        if self.uniformRayDensity and not given and fused and self._ext_variant is None:
            # one fused kernel: samples and (I, mJs, mJp) never leave the registers
            slot = self.shine_uniform_fused(bo, index_offset=index_offset, toGlobal=toGlobal)
            seeded = mcRays
            seededI = mcRays * self.xzE
            if bookkeeping:
                tmp_max = slot.view(torch.float64)
                if allreduce_max is not None:
                    tmp_max = allreduce_max(tmp_max)
                self.Imax = max(self.Imax, float(tmp_max))
                self.fluxConst = self.Imax * self.xzE
                bo.accepted = mcRays * self.fluxConst
                bo.acceptedE = float(bo.E.sum()) * self.fluxConst * SIE0
            bo.seeded = seeded
            bo.seededI = seededI
            bo.sourceWeight = self.xzE / self.nrays
            bo.parentId = getattr(self, 'uuid', self.name)
            return bo
        while True:
            s = samples if given else self.draw_samples(mcRays, dev, index_offset=index_offset)
            n = s['E'].numel()
            seeded += n
            gam = None
            if self.eEspread > 0:  # gamma += gamma * eEspread * normal (synchr.py:2053-2059)
                gam = self.gamma + (self.gamma * self.eEspread) * s['n5']
            I, mJs, mJp = self.build_I_map_device(s['E'], s['theta'], s['psi'], gamma=gam)
            if self.uniformRayDensity:
                seededI += n * self.xzE
                sourceWeight = self.xzE
                if bookkeeping:
                    tmp_max = I.max()
                    if allreduce_max is not None:
                        tmp_max = allreduce_max(tmp_max)
                    self.Imax = max(self.Imax, float(tmp_max))
                    self.fluxConst = self.Imax * self.xzE
                self.fill_beam(s, mJs, mJp, bo, toGlobal=toGlobal)
                length = n
                break
            # rejection sampling (sybase.py:1625-1641)
            tmp_max = I.max()
            if allreduce_max is not None:
                tmp_max = allreduce_max(tmp_max)
            seededI += float(I.sum()) * self.xzE
            sourceWeight = seededI / seeded
            if float(tmp_max) > self.Imax:
                self.Imax = float(tmp_max)
                self.fluxConst = self.Imax * self.xzE
            idx = torch.nonzero(self.Imax * s['u'] < I).flatten()
            npassed = min(int(idx.numel()), mcRays - length)
            if npassed > 0:
                self.fill_beam(s, mJs, mJp, bo, idx=idx, out_offset=length, n=npassed, toGlobal=toGlobal)
                length += npassed
            if length >= mcRays or given:
                break
        if length < mcRays:
            bo.filter_by_index(slice(0, length))
        if bookkeeping:
            bo.accepted = length * self.fluxConst
            bo.acceptedE = float(bo.E.sum()) * self.fluxConst * SIE0
        bo.seeded = seeded
        bo.seededI = seededI
        bo.sourceWeight = sourceWeight / self.nrays
        bo.parentId = getattr(self, 'uuid', self.name)
        return bo


def _intensities_on_mesh(self, energy, theta, psi, harmonic=None, eSpreadSigmas=3.5, eSpreadNSamples=36,
                         mode='constant'):
    """SourceBase.intensities_on_mesh (sources/sybase.py:740-928), resultKind='Stokes': [s0, s1/s0, s2/s0, s3/s0] in
    the shape (energy, theta, psi[, harmonic]); the field of every mesh point comes from the CUDA kernels through
    `build_I_map`, the incoherent averaging over energy spread and the angular convolution with the electron beam
    divergence (scipy.ndimage.gaussian_filter, like the reference) run on the small result arrays on the host."""
    if self.needReset:
        self.reset()
    energy, theta, psi = (np.asarray(a, dtype=float) for a in (energy, theta, psi))
    tomesh = [energy, theta, psi]
    iharmonic = None
    if harmonic is not None:
        tomesh.append(np.asarray(harmonic))
        iharmonic = len(tomesh)-1
    ispread = None
    if self.eEspread > 0:
        spr = np.linspace(-eSpreadSigmas, eSpreadSigmas, eSpreadNSamples)
        dgamma = self.gamma * spr * self.eEspread
        wspr = np.exp(-0.5 * spr**2)
        wspr /= wspr.sum()
        tomesh.append(dgamma)
        ispread = len(tomesh)-1
    mesh = np.meshgrid(*tomesh, indexing='ij')
    xE, xTheta, xPsi = mesh[0].ravel(), mesh[1].ravel(), mesh[2].ravel()
    sh = [len(energy), len(theta), len(psi)]
    xH = xG = None
    if iharmonic:
        xH = mesh[iharmonic].ravel()
        sh.append(len(harmonic))
    if ispread:
        xG = mesh[ispread].ravel()
        sh.append(len(dgamma))
    res = self.build_I_map(xE, xTheta, xPsi, xH, xG)
    Es = res[1].reshape(sh)
    Ep = res[2].reshape(sh)
    Is = (Es*np.conj(Es)).real.astype(float)
    Ip = (Ep*np.conj(Ep)).real.astype(float)
    Isp = Es*np.conj(Ep).astype(complex)
    if ispread:
        ws = wspr[(np.newaxis,) * (len(sh) - 1) + (slice(None),)]
        Is = (Is * ws).sum(axis=ispread)
        Ip = (Ip * ws).sum(axis=ispread)
        Isp = (Isp * ws).sum(axis=ispread)
    self.Is, self.Ip, self.Isp = Is, Ip, Isp
    s0 = Is + Ip
    s1 = Is - Ip
    s2 = 2. * np.real(Isp)
    s3 = -2. * np.imag(Isp)
    ss = [s0, s1, s2, s3]
    if (self.dxprime > 0 or self.dzprime > 0) and len(theta) > 1 and len(psi) > 1:
        from scipy.ndimage import gaussian_filter
        Sx = self.dxprime / (theta[1] - theta[0])
        Sz = self.dzprime / (psi[1] - psi[0])
        for ie in range(len(energy)):
            if harmonic is None:
                for arr in ss:
                    arr[ie, :, :] = gaussian_filter(arr[ie, :, :], [Sx, Sz], mode=mode)
            else:
                for ih in range(len(harmonic)):
                    for arr in ss:
                        arr[ie, :, :, ih] = gaussian_filter(arr[ie, :, :, ih], [Sx, Sz], mode=mode)
    with np.errstate(divide='ignore', invalid='ignore'):
        return [s0, np.where(s0, s1/s0, s0), np.where(s0, s2/s0, s0), np.where(s0, s3/s0, s0)]


def _tuning_curves(self, energy, theta, psi, harmonics, Ks):
    """Undulator.tuning_curves (sources/synchr.py:1710-1750): maximum flux of the given harmonics at the given K
    values through the aperture (theta, psi).  Returns (energy in keV, flux in ph/s/0.1%bw), rows = Ks, columns =
    harmonics."""
    try:
        dtheta, dpsi = theta[1] - theta[0], psi[1] - psi[0]
    except TypeError:
        dtheta, dpsi = 1, 1
    tunesE, tunesF = [], []
    tmpKy = self.Ky
    for K in Ks:
        self.Ky = K
        flux = None
        for e in energy:
            I0 = self.intensities_on_mesh([e], theta, psi, harmonics)[0]
            iflux = I0.sum(axis=(1, 2)) * dtheta * dpsi
            flux = iflux if flux is None else np.vstack((flux, iflux))
        argm = np.argmax(flux, axis=0)
        fluxm = np.max(flux, axis=0)
        tunesE.append(np.asarray(energy)[argm] / 1000.)
        tunesF.append(fluxm)
    self.Ky = tmpKy
    return np.array(tunesE).T, np.array(tunesF).T


def _power_vs_K(self, energy, theta, psi, harmonics, Ks):
    """Undulator.power_vs_K (sources/synchr.py:1752-1785): total power [W] through the aperture for each K."""
    try:
        dtheta, dpsi, dE = theta[1] - theta[0], psi[1] - psi[0], energy[1] - energy[0]
    except TypeError:
        dtheta, dpsi, dE = 1, 1, 1
    tmpKy = self.Ky
    powers = []
    for K in Ks:
        self.Ky = K
        I0 = self.intensities_on_mesh(energy, theta, psi, harmonics)[0]
        if self.distE == 'BW':
            I0 = I0 * 1e3
        else:
            I0 = I0 * np.asarray(energy)[(slice(None),) + (np.newaxis,) * (I0.ndim - 1)]
        powers.append(I0.sum() * dtheta * dpsi * dE * EV2ERG * 1e-7)
    self.Ky = tmpKy
    return np.array(powers)


Undulator.intensities_on_mesh = _intensities_on_mesh
Undulator.tuning_curves = _tuning_curves
Undulator.power_vs_K = _power_vs_K


class xrt_bm_params(ctypes.Structure):
    """struct xrt_bm_params of include/xrt_b200.h"""
    _fields_ = [('gamma0', ctypes.c_double), ('B', ctypes.c_double), ('ro', ctypes.c_double),
                ('K', ctypes.c_double), ('L0', ctypes.c_double), ('Np', ctypes.c_double), ('X0', ctypes.c_double),
                ('eI', ctypes.c_double), ('isMPW', ctypes.c_int), ('bw_mode', ctypes.c_int)]


class BendingMagnet(object):
    """Bending magnet with xrt's constructor, `build_I_map()` and `shine()` (sources/synchr.py:69-509; SourceBase
    sources/sybase.py:29-547), computing on the GPU: xrt_b200/csrc/source.cu (`bm_imap_kernel`, `bm_fill_kernel`,
    `bm_shine_uniform_kernel`).  Out of scope: filamentBeam, fixedEnergy, accuBeam."""
    isMPW = False

    def __init__(self, bl=None, name='BM', center=(0, 0, 0), nrays=100000, eE=6.0, eI=0.1, eEspread=0.,
                 eSigmaX=None, eSigmaZ=None, eEpsilonX=1., eEpsilonZ=0.01, betaX=9., betaZ=2., eMin=5000.,
                 eMax=15000., distE='eV', xPrimeMax=0.5, zPrimeMax=0.5, uniformRayDensity=False,
                 filamentBeam=False, pitch=0, yaw=0, B0=1., rho=None, targetOpenCL='CUDA',
                 precisionOpenCL='float64', seed=1234, **kwargs):
        if filamentBeam:
            raise NotImplementedError('filamentBeam is outside the hot path (SURVEY.md section 8f3)')
        if targetOpenCL is None:
            raise _lib.XrtB200Error('xrt_b200 sources always compute on the GPU (no numpy path)')
        self.bl = bl
        self.name = name
        if bl is not None and self not in bl.sources:
            bl.sources.append(self)
        self.center = list(center)
        self.nrays = int(nrays)
        self.pitch, self.yaw = pitch, yaw
        self.distE = distE
        self.uniformRayDensity = uniformRayDensity
        self.eE = float(eE)
        self.gamma = self.eE * 1e9 * EV2ERG / (M0 * C**2)  # sybase.py:155
        self.gamma2 = self.gamma**2
        self.eEspread = eEspread
        self.eI = float(eI)
        epsX = eEpsilonX * 1e-6 if eEpsilonX is not None else None
        epsZ = eEpsilonZ * 1e-6 if eEpsilonZ is not None else None

        def size(sigma, eps, beta):  # sybase.py:218-347
            if sigma is not None:
                d = sigma * 1e-3
            elif eps is not None and beta is not None:
                d = np.sqrt(eps * beta * 1e3)
            else:
                d = 0
            dprime = (eps / d if d > 0 else 0) if eps is not None else 0
            return d, dprime
        self.dx, self.dxprime = size(eSigmaX, epsX, betaX)
        self.dz, self.dzprime = size(eSigmaZ, epsZ, betaZ)
        self.eMin, self.eMax = float(eMin), float(eMax)
        self._xPrimeMax = Undulator._prime(xPrimeMax)
        self._zPrimeMax = Undulator._prime(zPrimeMax)
        E0 = SIE0 * C / 10
        if self.isMPW:  # synchr.py:88-94
            self.B = K2B * self.K / self.L0
            self.ro = M0 * C**2 * self.gamma / self.B / E0 / 1e6
            self.X0 = 0.5 * self.K * self.L0 / self.gamma / PI
            lim = self.K / self.gamma  # _xPrimeMaxAutoReduce (sybase.py:369-387), K0 = _Ky
            self._xPrimeMax = [np.sign(v) * min(abs(v), lim) for v in self._xPrimeMax]
        else:  # synchr.py:95-104
            self.Np = 0.5
            self.K = self.L0 = self.X0 = 0.
            self.B = B0
            self.ro = rho
            if self.ro:
                if not self.B:
                    self.B = M0 * C**2 * self.gamma / self.ro / E0 / 1e6
            elif self.B:
                self.ro = M0 * C**2 * self.gamma / self.B / E0 / 1e6
        self.seed = int(seed)
        self._draws = 0
        _lib.init(None)
        self.needReset = True

    def reset(self):  # sybase.py:479-547
        self.needReset = False
        self.Theta_min = float(self._xPrimeMax[0] - self.dxprime)
        self.Theta_max = float(self._xPrimeMax[1] + self.dxprime)
        self.Psi_min = float(self._zPrimeMax[0] - self.dzprime)
        self.Psi_max = float(self._zPrimeMax[1] + self.dzprime)
        self.E_min = float(min(self.eMin, self.eMax))
        self.E_max = float(max(self.eMin, self.eMax))
        self.Imax = 0.
        self.xzE = (self.E_max - self.E_min) * (self.Theta_max - self.Theta_min) * (self.Psi_max - self.Psi_min)
        self.fluxConst = self.Imax * self.xzE

    # ------------------------------------------------------------------ C structs
    def _bm_params(self):
        p = xrt_bm_params()
        p.gamma0, p.B, p.ro = float(self.gamma), float(self.B), float(self.ro)
        p.K, p.L0, p.Np, p.X0 = float(self.K), float(self.L0), float(self.Np), float(self.X0)
        p.eI = self.eI
        p.isMPW = int(self.isMPW)
        p.bw_mode = 1 if self.distE == 'BW' else 0
        return p

    _source_params = Undulator._source_params  # placement, sizes, pitch / yaw, azimuth: the same struct

    @property
    def E1(self):  # only read by the shared struct filler
        return 0.

    # ------------------------------------------------------------------ build_I_map
    def build_I_map(self, dde, ddtheta, ddpsi, harmonic=None, dg=None):
        """(I, Es, Ep) on host arrays like BendingMagnet.build_I_map (synchr.py:185-226).  *dg*: optional
        per-sample gamma offsets instead of the energy-spread draw."""
        if self.needReset:
            self.reset()
        lib = _lib.load()
        dde = np.ascontiguousarray(np.atleast_1d(dde), dtype=np.float64)
        n = len(dde)
        th = np.ascontiguousarray(np.asarray(ddtheta, dtype=np.float64) * np.ones(n))
        ps = np.ascontiguousarray(np.asarray(ddpsi, dtype=np.float64) * np.ones(n))
        gam = None
        if self.eEspread > 0 and n > 1:
            if dg is None:
                dg = np.random.normal(0, self.gamma*self.eEspread, dde.shape)
            gam = np.ascontiguousarray(self.gamma + dg)
        I = np.empty(n)
        Es = np.empty(n, dtype=np.complex128)
        Ep = np.empty(n, dtype=np.complex128)
        bm = self._bm_params()
        rc = lib.xrt_bm_imap_f64_host(n, ctypes.byref(bm), _lib.hptr(gam) if gam is not None else None,
                                      _lib.hptr(dde), _lib.hptr(th), _lib.hptr(ps), _lib.hptr(I), _lib.hptr(Es),
                                      _lib.hptr(Ep))
        _lib.check(rc, 'xrt_bm_imap_f64_host')
        return I, Es, Ep

    def build_I_map_device(self, w, theta, psi, gamma=None):
        lib = _lib.load()
        n = w.numel()
        I = torch.empty(n, dtype=F64, device=w.device)
        Es = torch.empty(n, dtype=C128, device=w.device)
        Ep = torch.empty(n, dtype=C128, device=w.device)
        bm = self._bm_params()
        rc = lib.xrt_bm_imap_f64(n, ctypes.byref(bm), gamma.data_ptr() if gamma is not None else None,
                                 w.data_ptr(), theta.data_ptr(), psi.data_ptr(), I.data_ptr(), Es.data_ptr(),
                                 Ep.data_ptr(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, 'xrt_bm_imap_f64')
        return I, Es, Ep

    # ------------------------------------------------------------------ sampling
    def draw_samples(self, n, device=None, index_offset=0):
        """Philox-sampled (E, theta, psi), the rejection uniform u, five standard normal deviates g1..g5, the
        energy-spread deviate g6 and the wiggler's period index k per sample (CUDA tensors)."""
        if self.needReset:
            self.reset()
        lib = _lib.load()
        dev = torch.device('cuda', torch.cuda.current_device()) if device is None else device
        names = ['E', 'theta', 'psi', 'g1', 'g2', 'g3', 'g4', 'u', 'g6', 'E2', 't2', 'p2', 'g5', 'h2', 'h3', 'h4',
                 'u2']
        buf = torch.empty((len(names), n), dtype=F64, device=dev)
        s = {k: buf[i] for i, k in enumerate(names)}
        window = (ctypes.c_double * 6)(self.E_min, self.E_max, self.Theta_min, self.Theta_max, self.Psi_min,
                                       self.Psi_max)
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        offset = int(index_offset) + self._draws
        self._draws += 1 << 40
        rc = lib.xrt_source_sample_f64(n, self.seed, offset, window, s['E'].data_ptr(), s['theta'].data_ptr(),
                                       s['psi'].data_ptr(), s['g1'].data_ptr(), s['g2'].data_ptr(),
                                       s['g3'].data_ptr(), s['g4'].data_ptr(), s['u'].data_ptr(), s['g6'].data_ptr(),
                                       stream)
        _lib.check(rc, 'xrt_source_sample_f64')
        # a second, disjoint counter range for the remaining deviates
        rc = lib.xrt_source_sample_f64(n, self.seed, offset + (1 << 39), window, s['E2'].data_ptr(),
                                       s['t2'].data_ptr(), s['p2'].data_ptr(), s['g5'].data_ptr(),
                                       s['h2'].data_ptr(), s['h3'].data_ptr(), s['h4'].data_ptr(),
                                       s['u2'].data_ptr(), None, stream)
        _lib.check(rc, 'xrt_source_sample_f64')
        if self.isMPW:
            s['k'] = torch.floor(s['u2'] * (2 * int(self.Np))) - int(self.Np)
        return s

    def fill_beam(self, s, mJs, mJp, out, idx=None, out_offset=0, n=None, toGlobal=True):
        lib = _lib.load()
        p = self._source_params(toGlobal)
        bm = self._bm_params()
        n = (len(idx) if idx is not None else s['E'].numel()) if n is None else n
        cout = out.c_struct()
        zeros = None

        def ptr(k):
            nonlocal zeros
            if k in s and s[k] is not None:
                return s[k].data_ptr()
            if zeros is None:
                zeros = torch.zeros_like(s['E'])
            return zeros.data_ptr()
        rc = lib.xrt_bm_fill_f64(
            n, ctypes.byref(p), ctypes.byref(bm), idx.data_ptr() if idx is not None else None, s['E'].data_ptr(),
            s['theta'].data_ptr(), s['psi'].data_ptr(), mJs.data_ptr(), mJp.data_ptr(), ptr('g1'), ptr('g2'),
            ptr('g3'), ptr('g4'), ptr('g5'), s['k'].data_ptr() if self.isMPW else None, ctypes.byref(cout),
            out_offset, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, 'xrt_bm_fill_f64')

    # ------------------------------------------------------------------ shine
    def shine(self, toGlobal=True, withAmplitudes=True, fixedEnergy=False, accuBeam=None, samples=None,
              index_offset=0, allreduce_max=None, fused=True):
        """Returns the source Beam (device resident) like BendingMagnet.shine (synchr.py:229-509).

        *samples* (optional): list of dicts (one per Monte-Carlo pass) of CUDA tensors E, theta, psi, u, g1..g5
        [, g6, k] to use instead of drawing them -- parity tests feed the deviates the numpy oracle drew; the
        accepted-ray deviates g1..g5, k are indexed like the samples (scattered by the test)."""
        if fixedEnergy or accuBeam is not None:
            raise NotImplementedError('fixedEnergy / accuBeam are outside the hot path (SURVEY 8f3)')
        if self.needReset:
            self.reset()
        lib = _lib.load()
        if self.uniformRayDensity:
            withAmplitudes = True
        mcRays = int(self.nrays * 1.2) if not self.uniformRayDensity else self.nrays
        dev = torch.device('cuda', torch.cuda.current_device())
        given = samples is not None
        nalloc = self.nrays if self.uniformRayDensity else self.nrays + mcRays
        bo = Beam(nalloc, withAmplitudes=withAmplitudes, device=dev, uninitialized=True)
        if self.uniformRayDensity and not given and fused:
            slot = torch.zeros(1, dtype=torch.int64, device=dev)
            p = self._source_params(toGlobal)
            bm = self._bm_params()
            cout = bo.c_struct()
            window = (ctypes.c_double * 6)(self.E_min, self.E_max, self.Theta_min, self.Theta_max, self.Psi_min,
                                           self.Psi_max)
            offset = int(index_offset) + self._draws
            self._draws += 1 << 40
            rc = lib.xrt_bm_shine_uniform_f64(len(bo), self.seed, offset, window, ctypes.byref(p), ctypes.byref(bm),
                                              ctypes.byref(cout), slot.data_ptr(),
                                              ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
            _lib.check(rc, 'xrt_bm_shine_uniform_f64')
            tmp_max = slot.view(torch.float64)
            if allreduce_max is not None:
                tmp_max = allreduce_max(tmp_max)
            self.Imax = max(self.Imax, float(tmp_max))
            self.fluxConst = self.Imax * self.xzE
            seeded, seededI, sourceWeight, length = mcRays, self.nrays * self.xzE, self.xzE, mcRays
        else:
            length = 0
            seeded = 0
            seededI = 0.
            ipass = 0
            while True:
                s = samples[ipass] if given else self.draw_samples(mcRays, dev, index_offset=index_offset)
                ipass += 1
                n = s['E'].numel()
                seeded += n
                gam = None
                if self.eEspread > 0:  # gamma += normal(0, gamma*eEspread) (synchr.py:192-197)
                    gam = self.gamma + (self.gamma * self.eEspread) * s['g6']
                I, mJs, mJp = self.build_I_map_device(s['E'], s['theta'], s['psi'], gamma=gam)
                if self.uniformRayDensity:
                    seededI += self.nrays * self.xzE
                    sourceWeight = self.xzE
                else:
                    seededI += float(I.sum()) * self.xzE
                    sourceWeight = seededI / seeded
                tmp_max = I.max()
                if allreduce_max is not None:
                    tmp_max = allreduce_max(tmp_max)
                if float(tmp_max) > self.Imax:
                    self.Imax = float(tmp_max)
                    self.fluxConst = self.Imax * self.xzE
                if self.uniformRayDensity:
                    self.fill_beam(s, mJs, mJp, bo, toGlobal=toGlobal)
                    length = n
                    break
                idx = torch.nonzero(self.Imax * s['u'] < I).flatten()
                npassed = int(idx.numel())
                if npassed > 0:
                    self.fill_beam(s, mJs, mJp, bo, idx=idx, out_offset=length, n=npassed, toGlobal=toGlobal)
                    length += npassed
                if length >= self.nrays or (given and ipass >= len(samples)):
                    break
        Esum = float(bo.E[:length].sum())  # the reference sums before cutting to nrays (synchr.py:482-490)
        if length != nalloc or length > self.nrays:
            bo.filter_by_index(slice(0, min(length, self.nrays)))
        if length >= self.nrays:
            bo.accepted = length * self.fluxConst
            bo.acceptedE = Esum * self.fluxConst * SIE0
            bo.seeded = seeded
            bo.seededI = seededI
            bo.sourceWeight = sourceWeight / self.nrays
        bo.parentId = getattr(self, 'uuid', self.name)
        return bo


BendingMagnet.intensities_on_mesh = _intensities_on_mesh


class Wiggler(BendingMagnet):
    """Multipole wiggler (sources/synchr.py:511-609): BendingMagnet with B from K and the period, the critical
    energy modulated along the horizontal angle and the source points spread over the 2 Np poles."""
    isMPW = True

    def __init__(self, *args, **kwargs):
        self.K = float(kwargs.pop('K', 8.446))
        self.L0 = float(kwargs.pop('period', 50))
        self.Np = kwargs.pop('n', 40)
        kwargs.setdefault('name', 'wiggler')
        super(Wiggler, self).__init__(*args, **kwargs)


XRT_GEOM_MAX_LINES = 16


class xrt_geom_params(ctypes.Structure):
    """struct xrt_geom_params of include/xrt_b200.h"""
    _fields_ = [('center', ctypes.c_double * 3), ('sinAz', ctypes.c_double), ('cosAz', ctypes.c_double),
                ('to_global', ctypes.c_int), ('n_rot', ctypes.c_int), ('rot_axis', ctypes.c_int * 3),
                ('rot_cos', ctypes.c_double * 3), ('rot_sin', ctypes.c_double * 3),
                ('dist', ctypes.c_int * 5), ('p0', ctypes.c_double * 5), ('p1', ctypes.c_double * 5),
                ('annulus_xz', ctypes.c_int), ('annulus_ac', ctypes.c_int),
                ('ann_xz', ctypes.c_double * 4), ('ann_ac', ctypes.c_double * 4), ('sqrt_pi2', ctypes.c_double),
                ('with_amp', ctypes.c_int),
                ('Jss', ctypes.c_double), ('Jpp', ctypes.c_double), ('Jsp_re', ctypes.c_double),
                ('Jsp_im', ctypes.c_double), ('Es_re', ctypes.c_double), ('Es_im', ctypes.c_double),
                ('Ep_re', ctypes.c_double), ('Ep_im', ctypes.c_double), ('ep_random', ctypes.c_int),
                ('distE', ctypes.c_int), ('e0', ctypes.c_double), ('e1', ctypes.c_double),
                ('n_lines', ctypes.c_int), ('weighted', ctypes.c_int),
                ('lines', ctypes.c_double * XRT_GEOM_MAX_LINES), ('cdf', ctypes.c_double * XRT_GEOM_MAX_LINES)]


def _auto_units_angle(angle, defaultFactor=1., aliases=None):
    """raycing.auto_units_angle (_flow_utils.py:74-115)."""
    if isinstance(angle, str):
        text = angle.strip()
        low = text.lower()
        if aliases is not None:
            for alias, value in aliases.items():
                if low.startswith(alias):
                    return value
        for unit, cut, factor in (('mrad', 'm', 1e-3), ('urad', 'u', 1e-6), ('nrad', 'n', 1e-9), ('rad', 'r', 1.)):
            if unit in angle:
                return float(angle.split(cut)[0].strip()) * factor
        if 'deg' in angle:
            return np.radians(float(angle.split('d')[0].strip()))
        return float(text) * defaultFactor
    if angle is None or isinstance(angle, (list, tuple)):
        return angle
    return angle * defaultFactor


class GeometricSource(object):
    """Geometric source with xrt's constructor and `shine()` (sources/geoms.py:193-534): ray origin, divergence and
    energy sampled with the given distribution laws ('normal', 'flat', 'annulus', None; distE 'normal', 'flat',
    'lines', None), every polarisation xrt knows, uniformRayDensity, canted sources, totalFlux.  One CUDA kernel
    draws the rays in HBM (Philox4x32-10 counters, so ray shards on several GPUs draw disjoint sub-streams), a second
    one normalises and places them: xrt_b200/csrc/source.cu.  Out of scope: filamentBeam, accuBeam."""

    def __init__(self, bl=None, name='', center=(0, 0, 0), nrays=100000, distx='normal', dx=0.32, disty=None, dy=0,
                 distz='normal', dz=0.018, distxprime='normal', dxprime=1e-3, distzprime='normal', dzprime=1e-4,
                 distE='lines', energies=(9000.,), energyWeights=None, polarization='horizontal',
                 filamentBeam=False, uniformRayDensity=False, pitch=0, roll=0, yaw=0, totalFlux=None, seed=1234,
                 targetOpenCL='CUDA', **kwargs):
        if filamentBeam:
            raise NotImplementedError('filamentBeam is outside the hot path (SURVEY.md section 8f3)')
        if targetOpenCL is None:
            raise _lib.XrtB200Error('xrt_b200 sources always compute on the GPU (no numpy path)')
        self.bl = bl
        self.name = name
        if bl is not None and self not in bl.sources:
            bl.sources.append(self)
        self.center = list(center)
        self.nrays = int(nrays)
        self.distx, self.dx = distx, dx
        self.disty, self.dy = disty, dy
        self.distz, self.dz = distz, dz
        self.distxprime, self.dxprime = distxprime, _auto_units_angle(dxprime)
        self.distzprime, self.dzprime = distzprime, _auto_units_angle(dzprime)
        self.distE, self.energies, self.energyWeights = distE, energies, energyWeights
        self.polarization = polarization
        self.uniformRayDensity = uniformRayDensity
        self.pitch, self.roll, self.yaw = (_auto_units_angle(v) for v in (pitch, roll, yaw))
        self.totalFlux = totalFlux
        self.seed = int(seed)
        self._draws = 0

    # ------------------------------------------------------------------ marshalling
    def _axis(self, p, k, dist, d):
        """_apply_distribution (geoms.py:370-405) -> (kind, p0, p1) of axis k."""
        if dist == 'normal':
            if self.uniformRayDensity:
                arr = np.atleast_1d(d)
                sigma = arr[0]
                cut = 5 * abs(sigma) if len(arr) < 2 else arr[-1]
                p.dist[k], p.p0[k], p.p1[k] = 3, float(sigma), float(cut)
            else:
                sigma = d[0] if isinstance(d, (list, tuple)) else d
                p.dist[k], p.p0[k] = 1, float(sigma)
        elif dist == 'flat':
            if isinstance(d, (list, tuple, np.ndarray)):
                aMin, aMax = d[0], d[1]
            else:
                if d <= 0:
                    return
                aMin, aMax = -d*0.5, d*0.5
            p.dist[k], p.p0[k], p.p1[k] = 2, float(aMin), float(aMax)

    def _pair(self, p, k1, k2, d1, v1, d2, v2, which):
        isAnnulus = False
        if d1 == 'annulus' or d2 == 'annulus':  # geoms.py:467-477
            isAnnulus = True
            if isinstance(v1, (list, tuple, np.ndarray)):
                rMin, rMax = v1
            else:
                isAnnulus = False
            if isinstance(v2, (list, tuple, np.ndarray)):
                phiMin, phiMax = v2
            else:
                phiMin, phiMax = 0, PI2
        if isAnnulus:
            setattr(p, 'annulus_' + which, 1)
            arr = getattr(p, 'ann_' + which)
            arr[0], arr[1], arr[2], arr[3] = float(rMin), float(rMax), float(phiMin), float(phiMax)
        else:
            self._axis(p, k1, d1, v1)
            self._axis(p, k2, d2, v2)

    def _polarization(self, p):
        """make_polarization (geoms.py:64-175)."""
        def fill(Jss, Jpp, Jsp, Es, Ep):
            p.Jss, p.Jpp, p.Jsp_re, p.Jsp_im = float(Jss), float(Jpp), complex(Jsp).real, complex(Jsp).imag
            p.Es_re, p.Es_im = complex(Es).real, complex(Es).imag
            if isinstance(Ep, str):
                p.ep_random = 1
            else:
                p.Ep_re, p.Ep_im = complex(Ep).real, complex(Ep).imag

        def linear(angle):
            Es, Ep = np.cos(angle), np.sin(angle)
            fill(Es*Es, Ep*Ep, Es*Ep, Es, Ep)

        def as_angle(value):
            if isinstance(value, (bool, np.bool_)):
                return None
            angle = _auto_units_angle(value, defaultFactor=np.pi / 180., aliases={'h': 0., 'v': np.pi / 2.})
            return angle if isinstance(angle, (int, float, np.integer, np.floating)) else None

        pol = self.polarization
        if pol is None:
            fill(0.5, 0.5, 0, 2**(-0.5), 'random phase')
        elif isinstance(pol, (tuple, list, np.ndarray)):
            if len(pol) != 4:
                raise ValueError('wrong coherency matrix: must be a 4-sequence!')
            p.Jss, p.Jpp, p.Jsp_re, p.Jsp_im = float(pol[0]), float(pol[1]), float(pol[2]), float(pol[3])
        elif isinstance(pol, str) and pol.lower().startswith('un'):
            fill(0.5, 0.5, 0, 2**(-0.5), 'random phase')
        elif isinstance(pol, str) and pol.lower().startswith('r'):
            fill(0.5, 0.5, 0.5j, 2**(-0.5), -1j * 2**(-0.5))
        elif isinstance(pol, str) and pol.lower().startswith('l'):
            fill(0.5, 0.5, -0.5j, 2**(-0.5), 1j * 2**(-0.5))
        else:
            angle = as_angle(pol.lower() if isinstance(pol, str) else pol)
            if angle is None:
                raise ValueError('wrong polarization!')
            linear(angle)

    def _energy(self, p):
        """make_energy (geoms.py:14-61)."""
        if self.distE is None:
            return
        eArr = np.atleast_1d(self.energies).astype(float)
        if self.distE == 'normal':
            eMean = eArr[0]
            eSigma = eArr[1] if len(eArr) == 2 else 0
            p.distE, p.e0, p.e1 = 1, float(eMean), float(0 if abs(eSigma) > 0.1*abs(eMean) else abs(eSigma))
        elif self.distE == 'flat':
            eMin = eArr[0]
            eMax = (eArr[1] or eArr[0]) if len(eArr) == 2 else eArr[0]
            p.distE, p.e0, p.e1 = 2, float(eMin), float(eMax)
        elif self.distE == 'lines':
            w = np.atleast_1d(self.energyWeights).astype(float) if self.energyWeights is not None else None
            if 0 in eArr:
                eArr = eArr[eArr > 0]
            if len(eArr) > XRT_GEOM_MAX_LINES:
                raise ValueError('at most {0} energy lines'.format(XRT_GEOM_MAX_LINES))
            p.distE, p.n_lines = 3, len(eArr)
            for i, e in enumerate(eArr):
                p.lines[i] = float(e)
            if w is not None and len(w) == len(eArr):
                p.weighted = 1
                cdf = w.cumsum()
                cdf /= cdf[-1]
                for i, c in enumerate(cdf):
                    p.cdf[i] = float(c)
        else:
            raise ValueError('unknown distE {0!r}'.format(self.distE))

    def _params(self, toGlobal, withAmplitudes):
        p = xrt_geom_params()
        for i in range(3):
            p.center[i] = float(self.center[i])
        p.sinAz = float(self.bl.sinAzimuth) if self.bl is not None else 0.
        p.cosAz = float(self.bl.cosAzimuth) if self.bl is not None else 1.
        p.to_global = int(bool(toGlobal))
        nrot = 0
        for ax, angle in ((2, self.yaw), (1, self.roll), (0, self.pitch)):  # rotate_beam default 'RzRyRx'
            if angle != 0:
                p.rot_axis[nrot], p.rot_cos[nrot], p.rot_sin[nrot] = ax, float(np.cos(angle)), float(np.sin(angle))
                nrot += 1
        p.n_rot = nrot
        p.sqrt_pi2 = PI2**0.5
        p.with_amp = int(bool(withAmplitudes))
        self._polarization(p)
        self._axis(p, 0, self.disty, self.dy)
        self._pair(p, 1, 2, self.distx, self.dx, self.distz, self.dz, 'xz')
        self._pair(p, 3, 4, self.distxprime, self.dxprime, self.distzprime, self.dzprime, 'ac')
        self._energy(p)
        return p

    # ------------------------------------------------------------------ shine
    def shine(self, toGlobal=True, withAmplitudes=False, accuBeam=None, samples=None, index_offset=0,
              allreduce_sum=None):
        """Returns the source Beam (device resident) like GeometricSource.shine (geoms.py:420-534).

        *samples* (optional): dict of CUDA tensors with the standard deviates to use instead of Philox draws, keys
        'pol', 'y', 'x', 'z', 'a', 'c', 'E' (the numpy oracle's record).  *allreduce_sum*: callable applied to the
        1-element tensors that are beam-wide (the a^2 + c^2 > 1 count, the intensity sum of totalFlux) when the
        rays are sharded over ranks."""
        if accuBeam is not None:
            raise NotImplementedError('accuBeam is outside the hot path')
        lib = _lib.load()
        _lib.init(None)
        if self.uniformRayDensity:
            withAmplitudes = True
        dev = torch.device('cuda', torch.cuda.current_device())
        bo = Beam(self.nrays, withAmplitudes=withAmplitudes, device=dev, uninitialized=True)
        p = self._params(toGlobal, withAmplitudes)
        flag = torch.zeros(1, dtype=torch.int64, device=dev)
        devs = (ctypes.c_void_p * 7)()
        if samples is not None:
            for k, key in enumerate(('pol', 'y', 'x', 'z', 'a', 'c', 'E')):
                t = samples.get(key)
                if t is not None:
                    devs[k] = t.data_ptr()
        cout = bo.c_struct()
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        offset = int(index_offset) + self._draws
        self._draws += 1 << 40
        args = (self.nrays, ctypes.byref(p), self.seed, offset, devs, ctypes.byref(cout), flag.data_ptr())
        _lib.check(lib.xrt_geometric_source_f64(*args, 1, stream), 'xrt_geometric_source_f64 (sample)')
        if allreduce_sum is not None:
            allreduce_sum(flag)
        if np.isscalar(self.totalFlux) and self.totalFlux > 0:  # make_flux_normalization (geoms.py:178-190)
            isum = (bo.Jss + bo.Jpp).sum().reshape(1)
            if allreduce_sum is not None:
                allreduce_sum(isum)
            intensitySum = float(isum)
            if intensitySum > 0:
                bo.sourceWeight = self.totalFlux / intensitySum
                bo.seeded = self.nrays
                bo.seededI = 1.
                bo.accepted = 1.
                bo.acceptedE = 1.
        _lib.check(lib.xrt_geometric_source_f64(*args, 2, stream), 'xrt_geometric_source_f64 (finish)')
        bo.parentId = getattr(self, 'uuid', self.name)
        return bo


class SourceFromField(Undulator):
    """Source based on a tabulated magnetic field (sources/synchr.py:612-1347): *customField* is an (N, 2 | 3 | 4)
    array [z (mm), {B_hor,} B_ver {, B_long}] (T), a text-file name or (fileName, loadtxt-kwargs).  The electron
    trajectory is integrated once on the host with the reference's RK4 scheme (it is sequential, O(field samples));
    the radiation integral along it for every (E, theta, psi) sample runs on the GPU (`custom_field_kernel`,
    xrt_b200/csrc/undulator_ext.cu).  Far field, or near field with *R0*.  Out of scope: filamentBeam, Excel input,
    the periodic test field of customField=None."""

    def __init__(self, *args, **kwargs):
        customField = kwargs.pop('customField', None)
        if customField is None:
            raise NotImplementedError('give customField (the built-in periodic test field is not on the hot path)')
        kwargs.setdefault('xPrimeMaxAutoReduce', False)
        kwargs.setdefault('zPrimeMaxAutoReduce', False)
        kwargs.setdefault('K', 1.)  # unused; the Undulator constructor wants some K
        R0 = kwargs.pop('R0', None)
        taper = kwargs.pop('taper', None)
        if taper is not None:
            raise ValueError('taper has no meaning for a tabulated field')
        Undulator.__init__(self, *args, **kwargs)
        self.R0 = R0
        if isinstance(customField, (tuple, list)):
            self.customFieldData = self.read_custom_field(customField[0], customField[1])
        elif isinstance(customField, np.ndarray):
            self.customFieldData = np.asarray(customField, dtype=float)
            z = self.customFieldData[:, 0]
            self.deviceLength = self._fwhm(z, np.abs(self.customFieldData[:, 1:]).max(axis=1))
        else:
            self.customFieldData = self.read_custom_field(customField, {})
        self.spl_kw = {'kind': 'cubic', 'bounds_error': False, 'fill_value': 'extrapolate'}
        self._traj = None
        self.needReset = True

    # ------------------------------------------------------------------ host: field, grid, trajectory
    def read_custom_field(self, fname, kwargs={}):  # synchr.py:695-705 (text files)
        if fname.endswith('.xls') or fname.endswith('.xlsx'):
            raise NotImplementedError('Excel field tables: convert to text or pass a numpy array')
        data = np.loadtxt(fname, **kwargs)
        self.deviceLength = self._fwhm(data[:, 0], np.abs(data[:, 1:]).max(axis=1))
        return data

    @staticmethod
    def _fwhm(z, a):  # synchr.py:707-710
        dz = z[1] - z[0]
        args = np.argwhere(a >= a.max()*0.5)
        return z[np.max(args)] - z[np.min(args)] + dz

    def report_E1(self):
        self.E1 = 0.

    def _magnetic_field(self, grid=None):  # synchr.py:727-765
        from scipy.interpolate import interp1d
        data = self.customFieldData
        dataz = data[:, 0]
        if grid is None:
            lenmm = np.abs(dataz[-1] - dataz[0])
            self.wtGrid = np.linspace(dataz[0], dataz[-1], int(lenmm*10))
            self.BGrid = np.linspace(dataz[0], dataz[-1], 2*len(self.wtGrid)-1)
            grid = self.BGrid
        cols = {2: (None, 1, None), 3: (1, 2, None), 4: (1, 2, 3)}.get(data.shape[1])
        if cols is None:
            raise ValueError('Unknown field table structure')
        out = []
        for c in cols:
            out.append(np.zeros_like(grid) if c is None else interp1d(dataz, data[:, c], **self.spl_kw)(grid))
        return out[0], out[1], out[2]

    def _build_integration_grid(self):  # synchr.py:988-1003
        self.tg_n, self.ag_n = clenshaw_curtis(self.quadm)
        dataz = self.customFieldData[:, 0]
        dstep = (dataz[-1] - dataz[0]) / float(self.gIntervals)
        dI = np.arange(0.5 * dstep + dataz[0], dataz[-1], dstep)
        self.tg = (dI[:, None]+0.5*dstep*self.tg_n).ravel()
        self.ag = (dI[:, None]*0+self.ag_n).ravel()
        self.dstep = dstep
        self._traj = None
        self._dev_tables = None
        self.gridSymmetric = False

    def _build_trajectory(self, Bx, By, Bz):
        """SourceFromField._build_trajectory_conv (synchr.py:1049-1147), not filamentBeam: RK4 for the transverse
        velocity in units of e/(m c gamma), three passes (mean velocity, mean position, tabulation)."""
        from scipy.interpolate import interp1d
        wt = self.wtGrid

        def f_beta(B, beta):
            return np.array((beta[1]*B[2]-B[1], B[0] - beta[0]*B[2]))

        def f_traj(beta):
            return np.array((beta[0], beta[1], -0.5*(beta[0]**2 + beta[1]**2)))

        def step(iB, h, beta, traj=None):
            B0, B1, B2 = ([Bx[iB+k], By[iB+k], Bz[iB+k]] for k in range(3))
            k1b = h * f_beta(B0, beta)
            k2b = h * f_beta(B1, beta + 0.5*k1b)
            k3b = h * f_beta(B1, beta + 0.5*k2b)
            k4b = h * f_beta(B2, beta + k3b)
            newb = beta + (k1b + 2*k2b + 2*k3b + k4b)/6.
            if traj is None:
                return newb
            k1t = h * f_traj(beta)
            k2t = h * f_traj(beta + 0.5*k1b)
            k3t = h * f_traj(beta + 0.5*k2b)
            k4t = h * f_traj(beta + k3b)
            return newb, traj + (k1t + 2*k2t + 2*k3t + k4t)/6.

        n = len(wt) - 1
        beta_next = np.zeros(2)
        beta0 = np.zeros(2)
        for i in range(n):
            h = wt[i+1] - wt[i]
            beta_next = step(2*i, h, beta_next)
            beta0 += h * beta_next
        beta0 /= -(wt[-1] - wt[0])
        beta_next = np.copy(beta0)
        traj_next = np.zeros(3)
        traj0 = np.zeros(3)
        betam_int = 0
        for i in range(n):
            h = wt[i+1] - wt[i]
            beta_next, traj_next = step(2*i, h, beta_next, traj_next)
            traj0 += h * traj_next
            betam_int += beta_next[0]**2 + beta_next[1]**2
        traj0 /= -(wt[-1] - wt[0])
        betam_int *= -0.5/n
        beta_next = np.copy(beta0)
        traj_next = np.copy(traj0)
        rows = [[beta0[0], beta0[1], traj0[0], traj0[1], traj0[2]]]
        for i in range(n):
            h = wt[i+1] - wt[i]
            beta_next, traj_next = step(2*i, h, beta_next, traj_next)
            rows.append([beta_next[0], beta_next[1], traj_next[0], traj_next[1], traj_next[2]])
        rows = np.array(rows)
        betax, betay, trajx, trajy, trajz = (interp1d(wt, rows[:, k], **self.spl_kw)(self.tg) for k in range(5))
        return betax, betay, betam_int, trajx, trajy, trajz

    def _field_tables(self):
        """(betam, [tg, ag, Bx, By, Bz, betax, betay, trajx, trajy, trajz]) on the integration grid."""
        if self._traj is None:
            Bx, By, Bz = self._magnetic_field()
            betax, betay, betam, trajx, trajy, trajz = self._build_trajectory(Bx, By, Bz)
            Bxt, Byt, Bzt = self._magnetic_field(self.tg)
            self._traj = (betam, [np.ascontiguousarray(a, dtype=np.float64) for a in
                                  (self.tg, self.ag, Bxt, Byt, Bzt, betax, betay, trajx, trajy, trajz)])
            self._dev_tables = None
        return self._traj

    def _tables(self, device):
        betam, tabs = self._field_tables()
        if self._dev_tables is None or self._dev_tables[0].device != device:
            self._dev_tables = [torch.from_numpy(a).to(device) for a in tabs]
        return self._dev_tables

    @property
    def _ext_variant(self):
        return 3, 0.  # never the fused far-field undulator kernel

    def _source_params(self, toGlobal):
        p = Undulator._source_params(self, toGlobal)
        # get_sigma_r02 = 2 CHeVcm / E * 10 * deviceLength / PI2**2 (synchr.py:712-725): L0 * Np -> deviceLength
        p.L0, p.Np = float(self.deviceLength), 1.
        p.eEspread = 0.  # SourceFromField.get_SIGMA has no energy-spread term
        return p

    # ------------------------------------------------------------------ build_I_map
    def _finish(self, xp, w, gamma, Is_local, Ip_local, betam):
        """synchr.py:1264-1272 / :1335-1347 on numpy (xp = np) or torch (xp = torch) arrays."""
        ab = 0.5 / np.pi / (1. - 0.5/gamma**2 + betam*EMC**2/gamma**2)
        bwFact = 0.001 if self.distE == 'BW' else 1./w
        Amp2Flux = FINE_STR * bwFact * self.eI / SIE0
        integralField = xp.abs(Is_local)**2 + xp.abs(Ip_local)**2
        sq = Amp2Flux**0.5
        return (Amp2Flux * 0.25 * self.dstep**2 * ab**2 * integralField,
                sq * Is_local * 0.5 * self.dstep * ab,
                sq * Ip_local * 0.5 * self.dstep * ab)

    def _integral_host(self, gamma, w, ddtheta, ddpsi):
        F = self.cl_precisionF
        betam, tabs = self._field_tables()
        NRAYS = len(w)
        scalarArgs = [np.int32(len(self.tg)), F(betam), F(self.R0 if self.R0 else 0.)]
        slicedRWArgs = [np.zeros(NRAYS, dtype=self.cl_precisionC), np.zeros(NRAYS, dtype=self.cl_precisionC)]
        return self.ucl.run_parallel('custom_field', scalarArgs, [F(gamma), F(w), F(ddtheta), F(ddpsi)], tabs,
                                     slicedRWArgs, None, NRAYS)

    def build_I_map(self, w, ddtheta, ddpsi, harmonic=None, dg=None):
        """(I, Es, Ep) on host arrays like SourceFromField.build_I_map (synchr.py:975-986, :1157-1272)."""
        if self.needReset:
            self.reset()
        w = np.atleast_1d(np.asarray(w, dtype=float))
        ddtheta = np.asarray(ddtheta, dtype=float) * np.ones_like(w)
        ddpsi = np.asarray(ddpsi, dtype=float) * np.ones_like(w)
        NRAYS = len(w)
        gamma = self.gamma
        if self.eEspread > 0:
            if dg is not None:
                gamma = gamma + dg
            else:
                gamma = gamma + gamma * self.eEspread * np.random.normal(size=NRAYS)
        gamma = gamma * np.ones(NRAYS)
        Is_local, Ip_local = self._integral_host(gamma, w, ddtheta, ddpsi)
        return self._finish(np, w, gamma, Is_local, Ip_local, self._field_tables()[0])

    def _convergence_value(self):
        w = np.array([self.E_max, self.E_max])
        th = np.array([self.Theta_max, self.Theta_max])
        ps = np.array([self.Psi_max, self.Psi_max])
        Is, Ip = self._integral_host(self.gamma * np.ones(2), w, th, ps)
        return np.abs(np.sqrt(np.abs(Is[0])**2 + np.abs(Ip[0])**2) * 0.5 * self.dstep)

    def build_I_map_device(self, w, theta, psi, gamma=None, out=None):
        if self.needReset:
            self.reset()
        lib = _lib.load()
        n = w.numel()
        tabs = self._tables(w.device)
        betam = self._field_tables()[0]
        if gamma is None:
            gamma = torch.full_like(w, self.gamma)
        raw = torch.empty((2, n), dtype=C128, device=w.device)
        rc = lib.xrt_custom_field_f64(n, len(self.tg), float(betam), float(self.R0 if self.R0 else 0.),
                                      gamma.data_ptr(), w.data_ptr(), theta.data_ptr(), psi.data_ptr(),
                                      *[t.data_ptr() for t in tabs], raw[0].data_ptr(), raw[1].data_ptr(),
                                      ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, 'xrt_custom_field_f64')
        return self._finish(torch, w, gamma, raw[0], raw[1], betam)
