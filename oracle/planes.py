"""Oracle: Screen.expose and RectangularAperture.propagate (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Restates xrt/backends/raycing/screens.py:225-302 (expose), apertures.py:333-413 (propagate) and
beamline.py:290-318 (xyz_from_xz) in numpy, in the reference's operation order."""
import numpy as np

from . import beam as ob
from .physconsts import CHBAR


def xyz_from_xz(sinAzimuth, cosAzimuth, x=None, z=None):
    if x is not None:
        norm = sum([xc**2 for xc in x])**0.5
        retx = [xc/norm for xc in x]
    else:
        retx = cosAzimuth, -sinAzimuth, 0.
    if z is not None:
        norm = sum([zc**2 for zc in z])**0.5
        retz = [zc/norm for zc in z]
    else:
        retz = 0., 0., 1.
    rety = np.cross(retz, retx)
    return [retx, rety, retz]


def expose(beam, center, xyz, lostNum, onlyPositivePath=False, compressX=None, compressZ=None, info=None):
    blo = ob.Beam(copyFrom=beam)
    ob.global_to_virgin_local(xyz, beam, blo, center)
    with np.errstate(divide='ignore', invalid='ignore'):
        path = -blo.y / blo.b
    condBad = np.isnan(path) | np.isinf(path)
    if onlyPositivePath:
        condBad = condBad | (path < 0)
    indBad = np.where(condBad)
    path[indBad] = 0.
    blo.state[indBad] = lostNum
    blo.path += path
    blo.x[:] += blo.a * path
    blo.z[:] += blo.c * path
    blo.y[:] = 0.
    if hasattr(blo, 'Es'):
        propPhase = np.exp(1e7j * (blo.E/CHBAR) * path)
        blo.Es *= propPhase
        blo.Ep *= propPhase
    good = (blo.state == 1) | (blo.state == 2)
    if info is not None and len(blo.state[good]) > 0:
        info['footprint'] = np.hstack((
            np.min(np.vstack((blo.x[good], blo.y[good], blo.z[good])), axis=1),
            np.max(np.vstack((blo.x[good], blo.y[good], blo.z[good])), axis=1))).reshape(2, 3)
    if compressX:
        blo.x[:] *= compressX
    if compressZ:
        blo.z[:] *= compressZ
    return blo


def propagate_rect(beam, center, xyz, blades, lostNum, isBeamStop=False, info=None):
    """`beam.state` is modified in place, as in the reference (apertures.py:364). Returns lo."""
    good = beam.state > 0
    lo = ob.Beam(copyFrom=beam)
    ob.global_to_virgin_local(xyz, beam, lo, center, good)
    path = -lo.y[good] / lo.b[good]
    lo.x[good] += lo.a[good] * path
    lo.z[good] += lo.c[good] * path
    lo.path[good] += path
    badIndices = np.zeros(len(beam.x), dtype=bool)
    for akind, d in blades.items():
        if akind.startswith('l'):
            badIndices[good] = badIndices[good] | (lo.x[good] < d)
        elif akind.startswith('r'):
            badIndices[good] = badIndices[good] | (lo.x[good] > d)
        elif akind.startswith('b'):
            badIndices[good] = badIndices[good] | (lo.z[good] < d)
        elif akind.startswith('t'):
            badIndices[good] = badIndices[good] | (lo.z[good] > d)
    if isBeamStop:
        badIndices[good] = np.invert(badIndices[good])
    beam.state[badIndices] = lostNum
    lo.state[:] = beam.state
    lo.y[good] = 0.
    if hasattr(lo, 'Es'):
        propPhase = np.exp(1e7j * (lo.E[good]/CHBAR) * path)
        lo.Es[good] *= propPhase
        lo.Ep[good] *= propPhase
    goodN = lo.state > 0
    if info is not None and goodN.sum() > 0:
        info['spotLimits'] = [lo.x[goodN].min(), lo.x[goodN].max(), lo.z[goodN].min(), lo.z[goodN].max()]
    return lo


def propagate_round(beam, center, xyz, r, lostNum, isBeamStop=False):
    """RoundAperture.propagate (apertures.py:769-825). `beam.state` is modified in place. Returns lo (with lo.r, the
    radii of the good rays)."""
    good = beam.state > 0
    lo = ob.Beam(copyFrom=beam)
    ob.global_to_virgin_local(xyz, beam, lo, center, good)
    path = -lo.y[good] / lo.b[good]
    lo.x[good] += lo.a[good] * path
    lo.z[good] += lo.c[good] * path
    lo.r = (lo.x[good]**2 + lo.z[good]**2)**0.5
    lo.path[good] += path
    badIndices = np.zeros(len(beam.x), dtype=bool)
    badIndices[good] = lo.r > r
    if isBeamStop:
        badIndices[good] = np.invert(badIndices[good])
    beam.state[badIndices] = lostNum
    lo.state[good] = beam.state[good]
    lo.y[good] = 0.
    if hasattr(lo, 'Es'):
        propPhase = np.exp(1e7j * (lo.E[good]/CHBAR) * path)
        lo.Es[good] *= propPhase
        lo.Ep[good] *= propPhase
    return lo


def expose_hemispheric(beam, center, xyz, R, lostNum, phiOffset=0, thetaOffset=0, onlyPositivePath=False):
    """HemisphericScreen.expose (screens.py:501-545); xyz = (x, y, z) local axes in the global system."""
    lx, ly, lz = xyz
    blo = ob.Beam(copyFrom=beam)
    sqb_2 = (beam.a * (beam.x-center[0]) + beam.b * (beam.y-center[1]) + beam.c * (beam.z-center[2]))
    sqc = ((beam.x-center[0])**2 + (beam.y-center[1])**2 + (beam.z-center[2])**2 - R**2)
    with np.errstate(invalid='ignore'):
        path = -sqb_2 + (sqb_2**2 - sqc)**0.5
    condBad = np.isnan(path) | np.isinf(path)
    if onlyPositivePath:
        condBad = condBad | (path < 0)
    indBad = np.where(condBad)
    path[indBad] = 0.
    blo.state[indBad] = lostNum
    blo.path += path
    rx = beam.x + beam.a*path - center[0]
    ry = beam.y + beam.b*path - center[1]
    rz = beam.z + beam.c*path - center[2]
    blo.z = rx*lz[0] + ry*lz[1] + rz*lz[2]
    blo.y = rx*ly[0] + ry*ly[1] + rz*ly[2]
    blo.x = rx*lx[0] + ry*lx[1] + rz*lx[2]
    blo.theta = np.arcsin(blo.z / R) - thetaOffset
    blo.phi = np.arctan2(blo.y, blo.x) - phiOffset
    if hasattr(blo, 'Es'):
        propPhase = np.exp(1e7j * (blo.E / CHBAR) * path)
        blo.Es *= propPhase
        blo.Ep *= propPhase
    return blo
