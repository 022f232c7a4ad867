"""Oracle: GeometricSource.shine (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Restates in numpy, in the reference's operation order and consuming the legacy global `np.random` stream in the
reference's order (so that after `np.random.seed(s)` the beam is reproduced bit for bit):
  * make_energy          xrt/backends/raycing/sources/geoms.py:14-61
  * make_polarization    geoms.py:64-175
  * make_flux_normalization  geoms.py:178-190
  * GeometricSource._apply_distribution / _set_annulus / shine   geoms.py:370-534 (no filamentBeam, no accuBeam)
`record` receives the standard deviates that were drawn (uniform [0, 1) or standard normal per stream), so that the
CUDA path can be fed the same numbers."""
import numpy as np

from . import beam as ob
from .physconsts import PI2


def is_sequence(a):
    return isinstance(a, (list, tuple, np.ndarray))


class GeomSpec(object):
    def __init__(self, **kw):
        self.center = [0., 0., 0.]
        self.azimuth = (0., 1.)
        self.nrays = 1000
        self.distx, self.dx = 'normal', 0.32
        self.disty, self.dy = None, 0
        self.distz, self.dz = 'normal', 0.018
        self.distxprime, self.dxprime = 'normal', 1e-3
        self.distzprime, self.dzprime = 'normal', 1e-4
        self.distE, self.energies, self.energyWeights = 'lines', (9000.,), None
        self.polarization = 'horizontal'
        self.uniformRayDensity = False
        self.pitch = self.roll = self.yaw = 0
        self.totalFlux = None
        self.__dict__.update(kw)


def _uniform(low, high, n, rec, key):
    u = np.random.random_sample(n)  # np.random.uniform(low, high, n) == low + (high - low) * random_sample
    rec[key] = u
    return low + (high - low) * u


def _normal(loc, scale, n, rec, key):
    g = np.random.standard_normal(n)  # np.random.normal(loc, scale, n) == loc + scale * gauss
    rec[key] = g
    return loc + scale * g


def make_energy(distE, energies, nrays, energyWeights, rec):
    eArr = np.atleast_1d(energies)
    if distE == 'normal':
        eMean = eArr[0]
        eSigma = eArr[1] if len(eArr) == 2 else 0
        E = _normal(eMean, 0 if abs(eSigma) > 0.1*abs(eMean) else abs(eSigma), nrays, rec, 'E')
    elif distE == 'flat':
        eMin = eArr[0]
        eMax = eArr[1] or eArr[0] if len(eArr) == 2 else eArr[0]
        E = _uniform(eMin, eMax, nrays, rec, 'E')
    elif distE == 'lines':
        eWeights = np.atleast_1d(energyWeights) if energyWeights is not None else None
        if 0 in eArr:
            eArr = eArr[eArr > 0]
        if energyWeights is not None and len(eArr) == len(eWeights):
            E = np.random.choice(eArr, size=nrays, p=eWeights)
        else:
            E = np.random.choice(eArr, size=nrays)
        rec['E'] = np.searchsorted(eArr, E).astype(float) if np.all(np.diff(eArr) > 0) else None
    return E


def make_polarization(polarization, bo, nrays, rec):
    def _fill_beam(Jss, Jpp, Jsp, Es, Ep):
        bo.Jss.fill(Jss)
        bo.Jpp.fill(Jpp)
        bo.Jsp.fill(Jsp)
        if hasattr(bo, 'Es'):
            bo.Es.fill(Es)
            if isinstance(Ep, str):
                u = np.random.uniform(size=int(nrays))
                rec['pol'] = u
                bo.Ep[:] = u * 2**(-0.5)
            else:
                bo.Ep.fill(Ep)

    def _fill_linear(angle):
        Es = np.cos(angle)
        Ep = np.sin(angle)
        _fill_beam(Es*Es, Ep*Ep, Es*Ep, Es, Ep)

    def _as_linear_angle(value):  # raycing.auto_units_angle with defaultFactor = pi/180 and the h / v aliases
        if isinstance(value, str):
            v = value.strip().lower()
            if v in ('h', 'horizontal'):
                return 0.
            if v in ('v', 'vertical'):
                return np.pi / 2.
            if v.endswith('rad'):
                return float(v[:-3])
            if v.endswith('deg'):
                return float(v[:-3]) * np.pi / 180.
            return float(v) * np.pi / 180.
        return value * np.pi / 180.

    if polarization is None:
        _fill_beam(0.5, 0.5, 0, 2**(-0.5), 'random phase')
    elif isinstance(polarization, (tuple, list, np.ndarray)):
        bo.Jss.fill(polarization[0])
        bo.Jpp.fill(polarization[1])
        bo.Jsp.fill(polarization[2] + 1j*polarization[3])
    elif isinstance(polarization, str):
        polarization = polarization.lower()
        if polarization.startswith('un'):
            _fill_beam(0.5, 0.5, 0, 2**(-0.5), 'random phase')
        elif polarization.startswith('r'):
            _fill_beam(0.5, 0.5, 0.5j, 2**(-0.5), -1j * 2**(-0.5))
        elif polarization.startswith('l'):
            _fill_beam(0.5, 0.5, -0.5j, 2**(-0.5), 1j * 2**(-0.5))
        else:
            _fill_linear(_as_linear_angle(polarization))
    else:
        _fill_linear(_as_linear_angle(polarization))


def _apply_distribution(src, axis, distaxis, daxis, bo, rec, key):
    n = src.nrays
    if distaxis == 'normal':
        if src.uniformRayDensity:
            daxisArr = np.atleast_1d(daxis)
            if len(daxisArr) < 2:
                sigma = daxisArr[0]
                cutLim = 5 * abs(sigma)
            else:
                sigma = daxisArr[0]
                cutLim = daxisArr[-1]
            axis[:] = _uniform(-cutLim, cutLim, n, rec, key)
            amp = np.exp(-axis**2 / sigma**2 / 2) / PI2**0.5 / sigma * 2 * cutLim
            bo.Jss *= amp
            bo.Jpp *= amp
            bo.Jsp *= amp
            amp = amp**0.5
            bo.Es *= amp
            bo.Ep *= amp
        else:
            sigma = daxis[0] if isinstance(daxis, (list, tuple)) else daxis
            axis[:] = _normal(0, sigma, n, rec, key)
    elif distaxis == 'flat':
        if is_sequence(daxis):
            aMin, aMax = daxis[0], daxis[1]
        else:
            if daxis <= 0:
                return
            aMin, aMax = -daxis*0.5, daxis*0.5
        axis[:] = _uniform(aMin, aMax, n, rec, key)


def _set_annulus(src, axis1, axis2, rMin, rMax, phiMin, phiMax, rec, key1, key2):
    if rMax > rMin:
        A = 2. / (rMax**2 - rMin**2)
        r = np.sqrt(2*_uniform(0, 1, src.nrays, rec, key1)/A + rMin**2)
    else:
        r = rMax
    phi = _uniform(phiMin, phiMax, src.nrays, rec, key2)
    axis1[:] = r * np.cos(phi)
    axis2[:] = r * np.sin(phi)


def _pair(src, bo, ax1, ax2, d1, v1, d2, v2, rec, k1, k2):
    isAnnulus = False
    if (d1 == 'annulus') or (d2 == 'annulus'):
        isAnnulus = True
        if is_sequence(v1):
            rMin, rMax = v1
        else:
            isAnnulus = False
        if is_sequence(v2):
            phiMin, phiMax = v2
        else:
            phiMin, phiMax = 0, PI2
    if isAnnulus:
        _set_annulus(src, ax1, ax2, rMin, rMax, phiMin, phiMax, rec, k1, k2)
    else:
        _apply_distribution(src, ax1, d1, v1, bo, rec, k1)
        _apply_distribution(src, ax2, d2, v2, bo, rec, k2)


def shine(src, toGlobal=True, withAmplitudes=False, record=None):
    rec = {} if record is None else record
    if src.uniformRayDensity:
        withAmplitudes = True
    bo = ob.Beam(src.nrays, withAmplitudes=withAmplitudes)
    bo.state[:] = 1
    make_polarization(src.polarization, bo, src.nrays, rec)
    _apply_distribution(src, bo.y, src.disty, src.dy, bo, rec, 'y')
    _pair(src, bo, bo.x, bo.z, src.distx, src.dx, src.distz, src.dz, rec, 'x', 'z')
    _pair(src, bo, bo.a, bo.c, src.distxprime, src.dxprime, src.distzprime, src.dzprime, rec, 'a', 'c')
    # make_flux_normalization (geoms.py:178-190)
    if np.isscalar(src.totalFlux) and src.totalFlux > 0:
        intensitySum = (bo.Jss + bo.Jpp).sum()
        if intensitySum > 0:
            bo.sourceWeight = src.totalFlux / intensitySum
            bo.seeded = len(bo.E)
            bo.seededI = 1.
            bo.accepted = 1.
            bo.acceptedE = 1.
    ac = bo.a**2 + bo.c**2
    if sum(ac > 1) > 0:
        bo.b[:] = (ac + 1)**0.5
        bo.a[:] /= bo.b
        bo.c[:] /= bo.b
        bo.b[:] = 1.0 / bo.b
    else:
        bo.b[:] = (1 - ac)**0.5
    if src.distE is not None:
        bo.E[:] = make_energy(src.distE, src.energies, src.nrays, src.energyWeights, rec)
    if src.pitch or src.roll or src.yaw:
        ob.rotate_beam(bo, pitch=src.pitch, roll=src.roll, yaw=src.yaw)
    if toGlobal:
        ob.virgin_local_to_global(src.azimuth, bo, src.center)
    return bo
