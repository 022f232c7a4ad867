"""Oracle: wave-mode bookkeeping around the Kirchhoff sum (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Restates, for a RectangularAperture as diffracting element and a flat Screen as receiving element,
  * Screen.prepare_wave              xrt/backends/raycing/screens.py:304-365
  * waves.prepare_wave               xrt/backends/raycing/waves.py:505-584
  * waves.diffract                   waves.py:606-831 (ConvexHull area :649-670, accumulation :707-717,
                                     direction de-phasing :719-737, flux normalisation :739-749, frames :756-790)
in numpy, in the reference's operation order."""
import numpy as np
from scipy.spatial import ConvexHull

from . import beam as ob, kirchhoff as ok, planes as op


class Plane(object):
    def __init__(self, center, sinAz, cosAz):
        self.center = list(center)
        self.sinAzimuth, self.cosAzimuth = sinAz, cosAz
        self.x, self.y, self.z = op.xyz_from_xz(sinAz, cosAz)

    def local_to_global(self, x=0, y=0, z=0):
        xglo = self.center[0] + x*self.x[0] + y*self.y[0] + z*self.z[0]
        yglo = self.center[1] + x*self.x[1] + y*self.y[1] + z*self.z[1]
        zglo = self.center[2] + x*self.x[2] + y*self.y[2] + z*self.z[2]
        return xglo, yglo, zglo


def prepare_wave(fromOE, wave, xglo, yglo, zglo):
    nrays = len(wave.x)
    wave.Es = np.zeros(nrays, dtype=complex)
    wave.Ep = np.zeros(nrays, dtype=complex)
    for k in ('EsAcc', 'EpAcc', 'aEacc', 'bEacc', 'cEacc'):
        setattr(wave, k, np.zeros_like(wave.Es))
    wave.Jss[:] = 0
    wave.Jpp[:] = 0
    wave.Jsp[:] = 0
    x, y, z = np.array(xglo), np.array(yglo), np.array(zglo)
    x -= fromOE.center[0]
    y -= fromOE.center[1]
    z -= fromOE.center[2]
    a0, b0 = fromOE.sinAzimuth, fromOE.cosAzimuth
    x[:], y[:] = ob.rotate_z(x, y, b0, a0)
    wave.xDiffr = x
    wave.yDiffr = y
    wave.zDiffr = z
    wave.rDiffr = (wave.xDiffr**2 + wave.yDiffr**2 + wave.zDiffr**2)**0.5
    wave.a[:] = wave.xDiffr / wave.rDiffr
    wave.b[:] = wave.yDiffr / wave.rDiffr
    wave.c[:] = wave.zDiffr / wave.rDiffr
    wave.path[:] = 0.
    wave.fromOE = fromOE
    wave.beamReflRays = np.int64(0)
    wave.beamReflSumJ = 0.
    wave.beamReflSumJnl = 0.
    wave.diffract_repeats = np.int64(0)
    return wave


def screen_prepare_wave(screen, prevOE, dim1, dim2, dy=0):
    d1s, d2s = np.meshgrid(dim1, dim2)
    d1s = d1s.flatten()
    d2s = d2s.flatten()
    dS = (dim1[1] - dim1[0]) * (dim2[1] - dim2[0])
    nrays = len(d1s)
    xglo, yglo, zglo = screen.local_to_global(x=d1s, z=d2s)
    wave = ob.Beam(nrays, withAmplitudes=True)
    wave.state[:] = 1
    wave.x[:] = d1s
    wave.y[:] = np.zeros_like(d1s) + dy
    wave.z[:] = d2s
    wave.dS = dS
    wave.toOE = screen
    wave.area = (np.ones_like(d1s) * dS).sum()
    return prepare_wave(prevOE, wave, xglo, yglo+dy, zglo)


def diffract(oeLocal, wave):
    oe = wave.fromOE
    good = oeLocal.state == 1
    goodlen = good.sum()
    if getattr(oeLocal, 'area', None) is None or oeLocal.area <= 0:
        impactPoints = np.vstack((oeLocal.x[good], oeLocal.z[good])).T
        hull = ConvexHull(impactPoints)
        outerPts = impactPoints[hull.vertices, :]
        lines = np.hstack([outerPts, np.roll(outerPts, -1, axis=0)])
        oeLocal.area = 0.5 * abs(sum(x1*y2-x2*y1 for x1, y1, x2, y2 in lines))
    n = [0, 1, 0]
    nl = oeLocal.a*n[0] + oeLocal.b*n[1] + oeLocal.c*n[2]
    wave.diffract_repeats += 1
    wave.beamReflRays += goodlen
    wave.beamReflSumJ += (oeLocal.Jss[good] + oeLocal.Jpp[good]).sum()
    wave.beamReflSumJnl += abs(((oeLocal.Jss[good] + oeLocal.Jpp[good]) * nl[good]).sum())
    Es, Ep, aE, bE, cE = ok.diffraction_integral(
        oeLocal.x[good], oeLocal.y[good], oeLocal.z[good], n, nl[good], oeLocal.Es[good], oeLocal.Ep[good],
        oeLocal.E[good], wave.xDiffr, wave.yDiffr, wave.zDiffr)
    wave.EsAcc += Es
    wave.EpAcc += Ep
    wave.aEacc += aE
    wave.bEacc += bE
    wave.cEacc += cE
    wave.E[:] = oeLocal.E[0]
    wave.Es[:] = wave.EsAcc
    wave.Ep[:] = wave.EpAcc
    wave.Jss[:] = (wave.Es * np.conj(wave.Es)).real
    wave.Jpp[:] = (wave.Ep * np.conj(wave.Ep)).real
    wave.Jsp[:] = wave.Es * np.conj(wave.Ep)
    toReal = np.exp(-1j * np.angle(wave.bEacc))
    wave.a[:] = (wave.aEacc * toReal).real
    wave.b[:] = (wave.bEacc * toReal).real
    wave.c[:] = (wave.cEacc * toReal).real
    norm = (wave.a**2 + wave.b**2 + wave.c**2)**0.5
    norm[norm == 0] = 1.
    wave.a /= norm
    wave.b /= norm
    wave.c /= norm
    norm = wave.dS * oeLocal.area * wave.beamReflSumJ
    de = wave.beamReflRays * wave.beamReflSumJnl * wave.diffract_repeats
    if de > 0:
        norm /= de
    else:
        norm = 0
    wave.Jss *= norm
    wave.Jpp *= norm
    wave.Jsp *= norm
    wave.Es *= norm**0.5
    wave.Ep *= norm**0.5
    glo = ob.Beam(copyFrom=wave)
    glo.x[:] = wave.xDiffr
    glo.y[:] = wave.yDiffr
    glo.z[:] = wave.zDiffr
    x, y, z = glo.x, glo.y, glo.z  # aperture local_to_global (apertures.py:446-458)
    xglo = oe.center[0] + x*oe.x[0] + y*oe.y[0] + z*oe.z[0]
    yglo = oe.center[1] + x*oe.x[1] + y*oe.y[1] + z*oe.z[1]
    zglo = oe.center[2] + x*oe.x[2] + y*oe.y[2] + z*oe.z[2]
    glo.x, glo.y, glo.z = xglo, yglo, zglo
    a, b, c = glo.a, glo.b, glo.c
    aglo = a*oe.x[0] + b*oe.y[0] + c*oe.z[0]
    bglo = a*oe.x[1] + b*oe.y[1] + c*oe.z[1]
    cglo = a*oe.x[2] + b*oe.y[2] + c*oe.z[2]
    glo.a, glo.b, glo.c = aglo, bglo, cglo
    toOE = wave.toOE
    wave.a[:], wave.b[:], wave.c[:] = glo.a, glo.b, glo.c
    wave.Jss[:], wave.Jpp[:], wave.Jsp[:] = glo.Jss, glo.Jpp, glo.Jsp
    wave.Es[:], wave.Ep[:] = glo.Es, glo.Ep
    a0, b0 = toOE.sinAzimuth, toOE.cosAzimuth
    wave.a[:], wave.b[:] = ob.rotate_z(wave.a, wave.b, b0, a0)
    return glo


# ---------------------------------------------------------------------------------------------- OE as the
# diffracting element (waves.py:675-686, :719-724, :756-767; OE.local_to_global oes/base.py:1165-1229)
def prepare_wave_from_oe(oe, wave, xglo, yglo, zglo):
    """waves.prepare_wave for fromOE with a rotationSequence (waves.py:547-572). `oe` is an oracle OESpec."""
    wave = prepare_wave(Plane(oe.center, oe.azimuth[0], oe.azimuth[1]), wave, xglo, yglo, zglo)
    x, y, z = wave.xDiffr, wave.yDiffr, wave.zDiffr
    hold = ob.Beam(len(x))
    hold.x, hold.y, hold.z = x, y, z
    ob.rotate_beam(hold, rotationSequence=oe.rotationSequence, pitch=-oe.pitch, roll=-oe.roll, yaw=-oe.yaw)
    if oe.extraPitch or oe.extraRoll or oe.extraYaw:
        ob.rotate_beam(hold, rotationSequence=oe.extraRotationSequence, pitch=-oe.extraPitch, roll=-oe.extraRoll,
                       yaw=-oe.extraYaw)
    wave.xDiffr, wave.yDiffr, wave.zDiffr = hold.x, hold.y, hold.z
    wave.rDiffr = (wave.xDiffr**2 + wave.yDiffr**2 + wave.zDiffr**2)**0.5
    wave.a[:] = wave.xDiffr / wave.rDiffr
    wave.b[:] = wave.yDiffr / wave.rDiffr
    wave.c[:] = wave.zDiffr / wave.rDiffr
    wave.fromOE = oe
    return wave


def screen_prepare_wave_from_oe(screen, oe, dim1, dim2):
    d1s, d2s = np.meshgrid(dim1, dim2)
    d1s = d1s.flatten()
    d2s = d2s.flatten()
    dS = (dim1[1] - dim1[0]) * (dim2[1] - dim2[0])
    xglo, yglo, zglo = screen.local_to_global(x=d1s, z=d2s)
    wave = ob.Beam(len(d1s), withAmplitudes=True)
    wave.state[:] = 1
    wave.x[:] = d1s
    wave.y[:] = np.zeros_like(d1s)
    wave.z[:] = d2s
    wave.dS = dS
    wave.toOE = screen
    wave.area = (np.ones_like(d1s) * dS).sum()
    return prepare_wave_from_oe(oe, wave, xglo, yglo, zglo)


def oe_local_to_global(oe, lb):
    """OE.local_to_global (oes/base.py:1165-1229) for a plain OE (no second crystal)."""
    if oe.dx:
        lb.x += oe.dx
    if oe.extraPitch or oe.extraRoll or oe.extraYaw:
        ob.rotate_beam(lb, rotationSequence='-'+oe.extraRotationSequence, pitch=oe.extraPitch, roll=oe.extraRoll,
                       yaw=oe.extraYaw)
    ob.rotate_beam(lb, rotationSequence='-'+oe.rotationSequence, pitch=oe.pitch, roll=oe.roll, yaw=oe.yaw)
    if oe.isParametric:
        s, phi, r = oe.xyz_to_param(lb.x, lb.y, lb.z)
        oeNormal = list(oe.local_n(s, phi))
    else:
        oeNormal = list(oe.local_n(lb.x, lb.y))
    roll = oe.roll + np.arctan2(oeNormal[-3], oeNormal[-1])
    lb.Jss[:], lb.Jpp[:], lb.Jsp[:] = ob.rotate_coherency_matrix(lb, slice(None), roll)
    if hasattr(lb, 'Es'):
        cosY, sinY = np.cos(roll), np.sin(roll)
        lb.Es[:], lb.Ep[:] = ob.rotate_y(lb.Es, lb.Ep, cosY, sinY)
    ob.virgin_local_to_global(oe.azimuth, lb, oe.center)


def diffract_from_oe(oeLocal, wave):
    """waves.diffract with an OE as wave.fromOE and a Screen as wave.toOE. All rays of oeLocal must be good (the
    reference's numpy integral broadcasts the un-masked normals against the masked samples, waves.py:840)."""
    oe = wave.fromOE
    good = oeLocal.state == 1
    goodlen = good.sum()
    if getattr(oeLocal, 'area', None) is None or oeLocal.area <= 0:
        impactPoints = np.vstack((oeLocal.x[good], oeLocal.y[good])).T
        hull = ConvexHull(impactPoints)
        outerPts = impactPoints[hull.vertices, :]
        lines = np.hstack([outerPts, np.roll(outerPts, -1, axis=0)])
        oeLocal.area = 0.5 * abs(sum(x1*y2-x2*y1 for x1, y1, x2, y2 in lines))
    if oe.isParametric:
        s, phi, r = oe.xyz_to_param(oeLocal.x, oeLocal.y, oeLocal.z)
        n = oe.local_n(s, phi)
    else:
        n = oe.local_n(oeLocal.x, oeLocal.y)[-3:]
    nl = (oeLocal.a*np.asarray([n[-3]]) + oeLocal.b*np.asarray([n[-2]]) + oeLocal.c*np.asarray([n[-1]])).flatten()
    wave.diffract_repeats += 1
    wave.beamReflRays += goodlen
    wave.beamReflSumJ += (oeLocal.Jss[good] + oeLocal.Jpp[good]).sum()
    wave.beamReflSumJnl += abs(((oeLocal.Jss[good] + oeLocal.Jpp[good]) * nl[good]).sum())
    nn = [np.asarray(n[i]) * np.ones(len(oeLocal.x)) for i in range(3)]
    Es, Ep, aE, bE, cE = ok.diffraction_integral(
        oeLocal.x[good], oeLocal.y[good], oeLocal.z[good], [nn[0][good], nn[1][good], nn[2][good]], nl[good],
        oeLocal.Es[good], oeLocal.Ep[good], oeLocal.E[good], wave.xDiffr, wave.yDiffr, wave.zDiffr)
    wave.EsAcc += Es
    wave.EpAcc += Ep
    wave.aEacc += aE
    wave.bEacc += bE
    wave.cEacc += cE
    wave.E[:] = oeLocal.E[0]
    wave.Es[:] = wave.EsAcc
    wave.Ep[:] = wave.EpAcc
    wave.Jss[:] = (wave.Es * np.conj(wave.Es)).real
    wave.Jpp[:] = (wave.Ep * np.conj(wave.Ep)).real
    wave.Jsp[:] = wave.Es * np.conj(wave.Ep)
    toRealComp = wave.cEacc if abs(wave.cEacc[0]) > abs(wave.bEacc[0]) else wave.bEacc
    toReal = np.exp(-1j * np.angle(toRealComp))
    wave.a[:] = (wave.aEacc * toReal).real
    wave.b[:] = (wave.bEacc * toReal).real
    wave.c[:] = (wave.cEacc * toReal).real
    norm = (wave.a**2 + wave.b**2 + wave.c**2)**0.5
    norm[norm == 0] = 1.
    wave.a /= norm
    wave.b /= norm
    wave.c /= norm
    norm = wave.dS * oeLocal.area * wave.beamReflSumJ
    de = wave.beamReflRays * wave.beamReflSumJnl * wave.diffract_repeats
    norm = norm / de if de > 0 else 0
    wave.Jss *= norm
    wave.Jpp *= norm
    wave.Jsp *= norm
    wave.Es *= norm**0.5
    wave.Ep *= norm**0.5
    glo = ob.Beam(copyFrom=wave)
    glo.x[:] = wave.xDiffr
    glo.y[:] = wave.yDiffr
    glo.z[:] = wave.zDiffr
    oe_local_to_global(oe, glo)
    # waves.py:776-784: the coherency rotation back to global acts on the local Es, Ep variables of diffract()
    # only (dead code for the outputs); then the wave takes the global beam and is turned by the azimuth
    toOE = wave.toOE
    wave.a[:], wave.b[:], wave.c[:] = glo.a, glo.b, glo.c
    wave.Jss[:], wave.Jpp[:], wave.Jsp[:] = glo.Jss, glo.Jpp, glo.Jsp
    wave.Es[:], wave.Ep[:] = glo.Es, glo.Ep
    a0, b0 = toOE.sinAzimuth, toOE.cosAzimuth
    wave.a[:], wave.b[:] = ob.rotate_z(wave.a, wave.b, b0, a0)
    return glo


# ---------------------------------------------------------------------------------------------- OE as the
# receiving element: OE.prepare_wave (oes/reflect.py:266-403) and the toOE frames of diffract (waves.py:786-824)
def oe_surface_z(oe, x, y):
    """z of the surface at local (x, y) as OE.prepare_wave takes it (reflect.py:335-340)."""
    if oe.isParametric:
        s, phi, r = oe.xyz_to_param(x, y, 0)
        r = oe.local_r(s, phi)
        return oe.param_to_xyz(s, phi, r)[2]
    return oe.local_z(x, y)


def oe_surface_centre_global(oe):
    """Middle point of the physical limits on the surface, global (reflect.py:344-361)."""
    cx = (oe.surfPhysX[1] + oe.surfPhysX[0])*0.5
    cy = (oe.surfPhysY[1] + oe.surfPhysY[0])*0.5
    lbc = ob.Beam(1)
    lbc.x[:] = cx
    lbc.y[:] = cy
    lbc.z[:] = oe_surface_z(oe, lbc.x.copy(), lbc.y.copy())
    oe_local_to_global(oe, lbc)
    return [lbc.x[0], lbc.y[0], lbc.z[0]]


def oe_prepare_wave(oe, prevOE, x, y, area='auto'):
    """OE.prepare_wave for sample positions (x, y) already drawn / meshed by the caller (reflect.py:302-333 is
    plain numpy on limPhys*).  `oe`, and `prevOE` when it is an OE, are OESpec; an aperture prevOE is a Plane."""
    from . import reflect as orf
    nsamples = len(x)
    if area == 'auto':
        area = (oe.surfPhysX[1] - oe.surfPhysX[0]) * (oe.surfPhysY[1] - oe.surfPhysY[0])
    lb = ob.Beam(nsamples, withAmplitudes=True)
    lb.state[:] = 1
    lb.x[:] = x
    lb.y[:] = y
    lb.z[:] = oe_surface_z(oe, x, y)
    oe_local_to_global(oe, lb)
    prevIsOE = hasattr(prevOE, 'rotationSequence')
    prevCenter = oe_surface_centre_global(prevOE) if prevIsOE else prevOE.center
    lb.a[:] = lb.x - prevCenter[0]
    lb.b[:] = lb.y - prevCenter[1]
    lb.c[:] = lb.z - prevCenter[2]
    norm = (lb.a**2 + lb.b**2 + lb.c**2)**0.5
    lb.a /= norm
    lb.b /= norm
    lb.c /= norm
    lb.x[:] = prevCenter[0]
    lb.y[:] = prevCenter[1]
    lb.z[:] = prevCenter[2]
    lbn = ob.Beam(1)
    lbn.b[:] = 0.
    lbn.c[:] = 1.
    oe_local_to_global(oe, lbn)
    a = lbn.x - prevCenter[0]
    b = lbn.y - prevCenter[1]
    c = lbn.z - prevCenter[2]
    norm = (a**2 + b**2 + c**2)**0.5
    areaNormalFact = abs(float(((a*lbn.a[0] + b*lbn.b[0] + c*lbn.c[0]) / norm)[0]))
    waveGlobal, waveLocal = orf.reflect(oe, lb)
    good = (waveLocal.state == 1) | (waveLocal.state == 2)
    waveGlobal.filter_by_index(good)
    waveLocal.filter_by_index(good)
    area *= good.sum() / float(len(good))
    waveLocal.area = area
    waveLocal.areaNormal = area * areaNormalFact
    waveLocal.dS = area / float(len(good))
    waveLocal.toOE = oe
    if prevIsOE:
        return prepare_wave_from_oe(prevOE, waveLocal, waveGlobal.x, waveGlobal.y, waveGlobal.z)
    return prepare_wave(prevOE, waveLocal, waveGlobal.x, waveGlobal.y, waveGlobal.z)


def to_oe_frames(wave, glo):
    """The part of diffract() that turns directions, coherency matrix and amplitudes into the local frame of a
    receiving OE and applies the obliquity factor (waves.py:786-824); call after diffract / diffract_from_oe."""
    toOE = wave.toOE
    if toOE.isParametric:
        s, phi, r = toOE.xyz_to_param(wave.x, wave.y, wave.z)
        oeNormal = list(toOE.local_n(s, phi))
    else:
        oeNormal = list(toOE.local_n(wave.x, wave.y))
    rollAngle = toOE.roll + np.arctan2(oeNormal[-3], oeNormal[-1])
    wave.Jss[:], wave.Jpp[:], wave.Jsp[:] = ob.rotate_coherency_matrix(wave, slice(None), -rollAngle)
    cosY, sinY = np.cos(rollAngle), np.sin(rollAngle)
    wave.Es[:], wave.Ep[:] = ob.rotate_y(wave.Es, wave.Ep, cosY, -sinY)
    hold = ob.Beam(len(wave.x))
    hold.x, hold.y, hold.z = wave.a, wave.b, wave.c
    ob.rotate_beam(hold, rotationSequence=toOE.rotationSequence, pitch=-toOE.pitch, roll=-toOE.roll, yaw=-toOE.yaw)
    if toOE.extraPitch or toOE.extraRoll or toOE.extraYaw:
        ob.rotate_beam(hold, rotationSequence=toOE.extraRotationSequence, pitch=-toOE.extraPitch,
                       roll=-toOE.extraRoll, yaw=-toOE.extraYaw)
    wave.a, wave.b, wave.c = hold.x, hold.y, hold.z
    norm = -wave.a*oeNormal[-3] - wave.b*oeNormal[-2] - wave.c*oeNormal[-1]
    norm = np.abs(norm)
    for b in (wave, glo):
        b.Jss *= norm
        b.Jpp *= norm
        b.Jsp *= norm
        b.Es *= norm**0.5
        b.Ep *= norm**0.5
    return glo
