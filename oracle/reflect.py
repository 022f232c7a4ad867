"""Oracle: OE.reflect / DCM.double_reflect (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Numpy restatement, in the reference's operation order, of
  * OE.reflect                       xrt/backends/raycing/oes/reflect.py:17-163
  * OE._reflect_local                oes/reflect.py:551-1139 (branches toWhere == 0: no material, 'mirror',
                                     'thin mirror', 'crystal' thick Bragg without mosaicity/TT)
  * OE._grating_deflection           oes/reflect.py:451-469, _get_asymmetric_reflection_grating :568-612
  * OE._bracketing / _set_t          oes/base.py:1231-1295
  * OE.find_intersection, find_dz    oes/base.py:801-885
  * OE._use_my_method / _use_Brent_method   oes/base.py:933-1048
  * OE.rays_good ('rect')            oes/base.py:1094-1163
  * surfaces: OE flat (base.py:675-742), ToroidMirror (oes/__init__.py:398-411),
              EllipticalMirrorParam (oes/parametric.py:213-246)
  * DCM.double_reflect               oes/dcm.py:247-354 (+ local_n2 :234-238)
"""
import numpy as np

from . import beam as ob
from .physconsts import CH, CHBAR

zEps = 1e-12
dt = 1e-5
maxIteration = 100
ds = 0.  # raycing.ds (raycing/__init__.py:90): margin used in multiple reflections
maxHalfSizeOfOE = 1000.
maxDepthOfOE = 100.


class OESpec(object):
    """Plain parameter bag (what the product passes to the CUDA kernel as OEParams)."""

    def __init__(self, **kw):
        self.azimuth = (0., 1.)  # (sinAzimuth, cosAzimuth)
        self.center = [0., 0., 0.]
        self.pitch = 0.
        self.roll = 0.  # roll + positionRoll
        self.yaw = 0.
        self.rotationSequence = 'RzRyRx'
        self.extraPitch = self.extraRoll = self.extraYaw = 0.
        self.extraRotationSequence = 'RzRyRx'
        self.dx = 0
        self.surface = 'flat'
        self.surfParams = {}
        self.surfPhysX = [-maxHalfSizeOfOE, maxHalfSizeOfOE]
        self.surfPhysY = [-maxHalfSizeOfOE, maxHalfSizeOfOE]
        self.surfOptX = None
        self.surfOptY = None
        self.overEdge = 'yMax'
        self.lostNum = -1
        self.material = None
        self.alpha = None
        self.isParametric = False
        self.gratingDensity = None  # ['x'|'y', rho0, c0, c1, ...] (base.py:213-230)
        self.order = 1
        self.__dict__.update(kw)

    sinAzimuth = property(lambda self: self.azimuth[0])
    cosAzimuth = property(lambda self: self.azimuth[1])

    # ------------------------------------------------------------------ surfaces
    def local_z(self, x, y):
        if self.surface == 'custom':  # a subclass surface (oracle/surfaces.py), like the reference's overrides
            return self.surfParams['obj'].local_z(x, y)
        if self.surface == 'flat':
            return np.zeros_like(y)
        if self.surface == 'toroid':
            R, r = self.surfParams['R'], self.surfParams['r']
            rx = 1 - (np.asarray(x)/r)**2
            rx[rx < 0] = 0.
            return y**2/2.0/R + r*(1 - rx**0.5)
        if self.surface == 'paraboloid':
            # ParaboloidFlatLens.local_z1 (refractive.py:394-399); ParabolicCylinderFlatLens passes x = 0 (:613-614)
            p = self.surfParams
            if p.get('isCylinder'):
                x = 0
            z = (x**2 + y**2) / (4 * p['focus'])
            if p['zmax'] is not None:
                z[z > p['zmax']] = p['zmax']
            return z
        raise ValueError(self.surface)

    def local_g(self, x, y, rho=-100.):
        """OE.local_g (base.py:688-717)."""
        rhoList = self.gratingDensity
        if rhoList is not None:
            coord = x if rhoList[0] == 'x' else y
            poly = 0.
            for ic, coeff in enumerate(rhoList[2:]):
                poly += (ic+1) * coeff * coord**ic
            N = rhoList[1] * poly
            if rhoList[0] == 'x':
                return N, np.zeros_like(N), np.zeros_like(N)
            elif rhoList[0] == 'y':
                return np.zeros_like(N), N, np.zeros_like(N)
        return 0, rho, 0

    def local_n(self, x, y):
        if self.surface == 'custom':
            return self.surfParams['obj'].local_n(x, y)
        if self.surface == 'flat':
            a = 0.
            b = 0.
            c = 1.
            if self.alpha:
                cosalpha, sinalpha = np.cos(self.alpha), np.sin(self.alpha)
                bAlpha, cAlpha = ob.rotate_x(b, c, cosalpha, -sinalpha)
                return [a, bAlpha, cAlpha, a, b, c]
            return [a, b, c]
        if self.surface == 'toroid':
            R, r = self.surfParams['R'], self.surfParams['r']
            rx = 1 - (np.asarray(x)/r)**2
            with np.errstate(divide='ignore', invalid='ignore'):
                ax = np.where(rx < 0, 0, rx**(-0.5))
            a = -x / r * ax
            b = -y / R
            c = 1.
            norm = (a**2 + b**2 + 1)**0.5
            return [a/norm, b/norm, c/norm]
        if self.surface == 'paraboloid':
            # ParaboloidFlatLens.local_n1 (refractive.py:405-419); cylinder: x = 0 (:616-617)
            p = self.surfParams
            if p.get('isCylinder'):
                x = 0
            a = -x / (2*p['focus'])
            b = -y / (2*p['focus'])
            if p['zmax'] is not None:
                z = (x**2 + y**2) / (4*p['focus'])
                if isinstance(a, np.ndarray):
                    a[z > p['zmax']] = 0
                if isinstance(b, np.ndarray):
                    b[z > p['zmax']] = 0
            c = np.ones_like(x)
            norm = (a**2 + b**2 + 1)**0.5
            return [a/norm, b/norm, c/norm]
        if self.surface == 'ellipse_param':
            p = self.surfParams
            s, phi = x, y
            A2s2 = np.array(p['ellipseA']**2 - s**2)
            A2s2[A2s2 <= 0] = 1e22
            nr = -p['ellipseB'] / p['ellipseA'] * s / np.sqrt(A2s2)
            norm = np.sqrt(nr**2 + 1)
            b = nr / norm
            if p['isCylindrical']:
                a = np.zeros_like(phi)
                c = 1. / norm
            else:
                a = -np.sin(phi) / norm
                c = -np.cos(phi) / norm
            bNew, cNew = ob.rotate_x(b, c, p['cosGamma'], -p['sinGamma'])
            return [a, bNew, cNew]
        raise ValueError(self.surface)

    def xyz_to_param(self, x, y, z):
        if self.surface == 'custom':
            return self.surfParams['obj'].xyz_to_param(x, y, z)
        p = self.surfParams
        yNew, zNew = ob.rotate_x(y - p['y0'], z - p['z0'], p['cosGamma'], p['sinGamma'])
        return yNew, np.arctan2(x, zNew), np.sqrt(x**2 + zNew**2)

    def param_to_xyz(self, s, phi, r):
        if self.surface == 'custom':
            return self.surfParams['obj'].param_to_xyz(s, phi, r)
        p = self.surfParams
        x = r * np.sin(phi)
        y = s
        z = r * np.cos(phi)
        yNew, zNew = ob.rotate_x(y, z, p['cosGamma'], -p['sinGamma'])
        return x, yNew + p['y0'], zNew + p['z0']

    def local_r(self, s, phi):
        if self.surface == 'custom':
            return self.surfParams['obj'].local_r(s, phi)
        p = self.surfParams
        r = p['ellipseB'] * np.sqrt(abs(1 - s**2 / p['ellipseA']**2))
        if p['isCylindrical']:
            r /= abs(np.cos(phi))
        return np.where(abs(phi) > np.pi/2, r, np.ones_like(phi)*1e20)


# ---------------------------------------------------------------------- intersection machinery
def find_dz(oe, local_f, t, x0, y0, z0, a, b, c, invertNormal, derivOrder=0):
    x = x0 + a*t
    y = y0 + b*t
    z = z0 + c*t
    if derivOrder == 0:
        if oe.isParametric:
            if local_f is None:
                local_f = oe.local_r
            diffSign = -1
        else:
            if local_f is None:
                local_f = oe.local_z
            diffSign = 1
    else:  # base.py:818-820
        if local_f is None:
            local_f = oe.local_n
    if oe.isParametric:
        x, y, z = oe.xyz_to_param(x, y, z)
    surf = local_f(x, y)
    fe = getattr(oe, 'figureError', None)
    if fe is not None and derivOrder == 0:  # base.py:825-829
        if oe.isParametric:  # local_r_distorted (base.py:790-797): the map acts along z at the Cartesian image of (s, phi)
            r0 = oe.local_r(x, y)
            X, Y, Z = oe.param_to_xyz(x, y, r0)
            Z += fe.local_z_distorted(X, Y)
            _, _, r1 = oe.xyz_to_param(X, Y, Z)
            surf = surf + (r1 - r0)
        else:
            surf = surf + fe.local_z_distorted(x, y)
    if derivOrder == 0:
        ind = np.isnan(surf)
        if ind.sum() > 0:
            surf[ind] = 0
        dz = (z - surf) * diffSign * invertNormal
    else:  # base.py:842-845: the ray direction along the surface normal
        dz = (a*surf[-3] + b*surf[-2] + c*surf[-1]) * invertNormal
    return dz, x, y, z


def _use_my_method(oe, local_f, t1, t2, x, y, z, a, b, c, invertNormal, dz1, dz2, tMin, tMax, x2, y2, z2, ind,
                   derivOrder=0):
    numit = 2
    while (ind.sum() > 0) and (numit < maxIteration):
        t = t1[ind]
        dz = dz1[ind]
        t1[ind] = t2[ind]
        dz1[ind] = dz2[ind]
        t2[ind] = t - (t1[ind]-t) * dz / (dz1[ind]-dz)
        swap = t2[ind] < tMin
        t2[np.where(ind)[0][swap]] = tMin
        swap = t2[ind] > tMax
        t2[np.where(ind)[0][swap]] = tMax
        dz2[ind], x2[ind], y2[ind], z2[ind] = find_dz(
            oe, local_f, t2[ind], x[ind], y[ind], z[ind], a[ind], b[ind], c[ind], invertNormal, derivOrder)
        swap = np.sign(dz2[ind]) == np.sign(dz1[ind])
        t1[np.where(ind)[0][swap]] = t[swap]
        dz1[np.where(ind)[0][swap]] = dz[swap]
        ind = ind & (abs(dz2) > zEps)
        numit += 1
    return t2, x2, y2, z2, numit


def _use_Brent_method(oe, local_f, t1, t2, x, y, z, a, b, c, invertNormal, dz1, dz2, tMin, tMax, x2, y2, z2, ind,
                      derivOrder=0):
    swap = abs(dz1[ind]) < abs(dz2[ind])
    if swap.sum() > 0:
        t1[np.where(ind)[0][swap]], t2[np.where(ind)[0][swap]] =\
            t2[np.where(ind)[0][swap]], t1[np.where(ind)[0][swap]]
        dz1[np.where(ind)[0][swap]], dz2[np.where(ind)[0][swap]] =\
            dz2[np.where(ind)[0][swap]], dz1[np.where(ind)[0][swap]]
    t3 = np.copy(t1)
    dz3 = np.copy(dz1)
    t4 = np.zeros_like(t1)
    mflag = np.ones_like(t1, dtype='bool')
    numit = 2
    ind = ind & (abs(dz2) > zEps)
    while (ind.sum() > 0) and (numit < maxIteration):
        xa, xb, xc, xd = t1[ind], t2[ind], t3[ind], t4[ind]
        fa, fb, fc = dz1[ind], dz2[ind], dz3[ind]
        mf = mflag[ind]
        xs = np.empty_like(xa)
        inq = (fa != fc) & (fb != fc)
        if inq.sum() > 0:
            xai = xa[inq]
            xbi = xb[inq]
            xci = xc[inq]
            fai = fa[inq]
            fbi = fb[inq]
            fci = fc[inq]
            xs[inq] = \
                xai * fbi * fci / (fai-fbi) / (fai-fci) + \
                fai * xbi * fci / (fbi-fai) / (fbi-fci) + \
                fai * fbi * xci / (fci-fai) / (fci-fbi)
        inx = ~inq
        if inx.sum() > 0:
            xai = xa[inx]
            xbi = xb[inx]
            fai = fa[inx]
            fbi = fb[inx]
            xs[inx] = xbi - fbi * (xbi-xai) / (fbi-fai)
        cond1 = ((xs < (3*xa + xb) / 4.) & (xs < xb) | (xs > (3*xa + xb) / 4.) & (xs > xb))
        cond2 = mf & (abs(xs - xb) >= (abs(xb - xc) / 2.))
        cond3 = (~mf) & (abs(xs - xb) >= (abs(xc - xd) / 2.))
        cond4 = mf & (abs(xb - xc) < zEps)
        cond5 = (~mf) & (abs(xc - xd) < zEps)
        conds = cond1 | cond2 | cond3 | cond4 | cond5
        xs[conds] = (xa[conds] + xb[conds]) / 2.
        mf = conds
        fs, x2[ind], y2[ind], z2[ind] = find_dz(
            oe, local_f, xs, x[ind], y[ind], z[ind], a[ind], b[ind], c[ind], invertNormal, derivOrder)
        xd[:] = xc[:]
        xc[:] = xb[:]
        fc[:] = fb[:]
        fafsNeg = ((fa < 0) & (fs > 0)) | ((fa > 0) & (fs < 0))
        xb[fafsNeg] = xs[fafsNeg]
        fb[fafsNeg] = fs[fafsNeg]
        fafsPos = ~fafsNeg
        xa[fafsPos] = xs[fafsPos]
        fa[fafsPos] = fs[fafsPos]
        swap = abs(fa) < abs(fb)
        xa[swap], xb[swap] = xb[swap], xa[swap]
        fa[swap], fb[swap] = fb[swap], fa[swap]
        t1[ind], t2[ind], t3[ind], t4[ind] = xa, xb, xc, xd
        dz1[ind], dz2[ind], dz3[ind] = fa, fb, fc
        mflag[ind] = mf
        ind = ind & (abs(dz2) > zEps)
        numit += 1
    return t2, x2, y2, z2, numit


def find_intersection(oe, local_f, t1, t2, x, y, z, a, b, c, invertNormal, info=None, derivOrder=0):
    dz1, x1, y1, z1 = find_dz(oe, local_f, t1, x, y, z, a, b, c, invertNormal, derivOrder)
    dz2, x2, y2, z2 = find_dz(oe, local_f, t2, x, y, z, a, b, c, invertNormal, derivOrder)
    tMin = t1.min()
    tMax = t2.max()
    ind1 = dz1 <= 0
    ind2 = dz2 >= 0
    dz2[ind1 | ind2] = 0
    t2[ind1] = t1[ind1]
    x2[ind1] = x1[ind1]
    y2[ind1] = y1[ind1]
    z2[ind1] = z1[ind1]
    ind = ~(ind1 | ind2)
    useBrent = abs(dz2).max() > abs(dz1).max()*20
    if info is not None:
        info.update(tMin=tMin, tMax=tMax, useBrent=bool(useBrent), maxdz1=abs(dz1).max(), maxdz2=abs(dz2).max())
    method = _use_Brent_method if useBrent else _use_my_method
    t2, x2, y2, z2, numit = method(oe, local_f, t1, t2, x, y, z, a, b, c, invertNormal, dz1, dz2, tMin, tMax,
                                   x2, y2, z2, ind, derivOrder)
    if info is not None:
        info['numit'] = numit
    return t2, x2, y2, z2, ind1


def _set_t(xyz, abc, surfPhys=None, defSize=maxHalfSizeOfOE):
    if surfPhys is None:
        limMin = -defSize
        limMax = defSize
    else:
        limMin = surfPhys[0] if surfPhys[0] > -np.inf else -defSize
        limMax = surfPhys[1] if surfPhys[1] < np.inf else defSize
    if abc[0] > 0:
        tMin = (limMin-xyz)/abc - dt
        tMax = (limMax-xyz)/abc + dt
    else:
        tMin = (limMax-xyz)/abc - dt
        tMax = (limMin-xyz)/abc + dt
    return tMin, tMax


def _bracketing(oe, x, y, z, a, b, c, surfPhysX, surfPhysY, mainPart, info=None, isMulti=False, local_n=None,
                invertNormal=1, needElevationMap=False):
    try:
        maxa = np.max(abs(a[mainPart]))
        maxb = np.max(abs(b[mainPart]))
        maxc = np.max(abs(c[mainPart]))
    except ValueError:
        maxa, maxb, maxc = 0, 1, 0
    maxMax = max(maxa, maxb, maxc)
    if maxMax == maxa:
        tMin, tMax = _set_t(x, a, surfPhysX)
        axis = 0
    elif maxMax == maxb:
        tMin, tMax = _set_t(y, b, surfPhysY)
        axis = 1
    else:
        tMin, tMax = _set_t(z, c, defSize=maxDepthOfOE)
        axis = 2
    if info is not None:
        info.update(axis=axis, maxa=maxa, maxb=maxb, maxc=maxc)
    tMin[tMin < -1e6*zEps] = -1e6*zEps
    if isMulti:  # base.py:1279-1289: leave the surface first -- the root of the direction along the normal
        tMin[:] = 0
        tMaxTmp = np.copy(tMax)
        infoD = {} if info is not None else None
        tMax, _, _, _, _ = find_intersection(oe, local_n, tMin, tMax, x, y, z, a, b, c, invertNormal, infoD,
                                             derivOrder=1)
        if info is not None:
            info['deriv'] = infoD
        if needElevationMap:  # base.py:1285-1287: (dz, x, y, z) at the derivative root = the highest point
            info_e = find_dz(oe, None, tMax, x, y, z, a, b, c, invertNormal)
            if info is not None:
                info['elevation'] = info_e
            elevation = info_e
        else:
            elevation = None
        tMin = tMax + ds
        tMax = tMaxTmp
        if needElevationMap:
            return tMin, tMax, elevation
    if needElevationMap:
        return tMin, tMax, None
    return tMin, tMax


def rays_good(oe, x, y, surfPhysX, surfPhysY, surfOptX, surfOptY):
    locState = np.ones(x.size, dtype=np.int32)
    if str(getattr(oe, 'shape', 'rect')).startswith('ro'):  # base.py:1140-1156
        centerX = (surfPhysX[0]+surfPhysX[1]) * 0.5
        if np.isnan(centerX):
            centerX = 0
        radiusX = (surfPhysX[1]-surfPhysX[0]) * 0.5
        centerY = (surfPhysY[0]+surfPhysY[1]) * 0.5
        radiusY = (surfPhysY[1]-surfPhysY[0]) * 0.5
        if np.isnan(centerY):
            centerY = 0
        if not np.isinf(radiusX):
            locState[((x-centerX)/radiusX)**2 + ((y-centerY)/radiusY)**2 > 1] = oe.lostNum
        return locState
    if surfOptX is not None:
        locState[((surfPhysX[0] <= x) & (x < surfOptX[0])) | ((surfOptX[1] <= x) & (x < surfPhysX[1]))] = 2
    if surfOptY is not None:
        locState[((surfPhysY[0] <= y) & (y < surfOptY[0])) | ((surfOptY[1] <= y) & (y < surfPhysY[1]))] = 2
    ovE = str(oe.overEdge).lower()
    outside = (x < surfPhysX[0]) | (x > surfPhysX[1]) | (y < surfPhysY[0]) | (y > surfPhysY[1])
    over = np.zeros_like(outside)
    if 'xmin' in ovE:
        over |= x < surfPhysX[0]
    if 'xmax' in ovE:
        over |= x > surfPhysX[1]
    if 'ymin' in ovE:
        over |= y < surfPhysY[0]
    if 'ymax' in ovE:
        over |= y > surfPhysY[1]
    locState[outside] = oe.lostNum
    locState[over] = 3
    return locState


def _grating_deflection(goodN, lb, g, oeNormal, beamInDotNormal, order=1, sig=None):
    beamInDotG = lb.a[goodN]*g[0] + lb.b[goodN]*g[1] + lb.c[goodN]*g[2]
    G2 = g[0]**2 + g[1]**2 + g[2]**2
    locOrder = order
    lb.order = np.zeros(len(lb.a))  # reflect.py:458-459
    lb.order[goodN] = locOrder
    orderLambda = locOrder * CH / lb.E[goodN] * 1e-7
    u = beamInDotNormal**2 - 2*beamInDotG*orderLambda - G2*orderLambda**2
    gs = np.sign(beamInDotNormal) if sig is None else sig
    dn = beamInDotNormal + gs*np.sqrt(abs(u))
    a_out = lb.a[goodN] - oeNormal[-3]*dn + g[0]*orderLambda
    b_out = lb.b[goodN] - oeNormal[-2]*dn + g[1]*orderLambda
    c_out = lb.c[goodN] - oeNormal[-1]*dn + g[2]*orderLambda
    norm = (a_out**2 + b_out**2 + c_out**2)**0.5
    return a_out/norm, b_out/norm, c_out/norm


def reflect_local(oe, good, lb, vlb, pitch, roll, yaw, dx=None, dy=None, dz=None, local_n=None, material=None,
                  is2ndXtal=False, surfPhysX=None, surfPhysY=None, surfOptX=None, surfOptY=None, info=None,
                  noIntersectionSearch=False, fromVacuum=True, isMulti=False, needElevationMap=False):
    if local_n is None:
        local_n = oe.local_n
    if surfPhysX is None:
        surfPhysX, surfPhysY, surfOptX, surfOptY = oe.surfPhysX, oe.surfPhysY, oe.surfOptX, oe.surfOptY
    extraAnglesSign = 1.
    if is2ndXtal:
        ob.rotate_beam(lb, good, roll=-np.pi)
        extraAnglesSign = -1.
    ob.rotate_beam(lb, good, rotationSequence=oe.rotationSequence, pitch=-pitch, roll=-roll, yaw=-yaw)
    if oe.extraPitch or oe.extraRoll or oe.extraYaw:
        ob.rotate_beam(lb, good, rotationSequence=oe.extraRotationSequence,
                       pitch=-extraAnglesSign*oe.extraPitch, roll=-oe.extraRoll, yaw=-extraAnglesSign*oe.extraYaw)
    if dx:
        lb.x[good] -= dx
    if dy:
        lb.y[good] -= dy
    if dz:
        lb.z[good] -= dz
    invertNormal = 1 if fromVacuum else -1  # reflect.py:638-641
    mainPartForBracketing = lb.state[good] == 1
    tMin = np.zeros_like(lb.x)
    tMax = np.zeros_like(lb.x)
    if needElevationMap:  # reflect.py:647-659
        tMin[good], tMax[good], elev = _bracketing(
            oe, lb.x[good], lb.y[good], lb.z[good], lb.a[good], lb.b[good], lb.c[good], surfPhysX, surfPhysY,
            mainPartForBracketing, info, isMulti=isMulti, local_n=local_n, invertNormal=invertNormal,
            needElevationMap=True)
        if elev:
            lb.elevationD[good] = elev[0]
            if oe.isParametric:
                tX, tY, tZ = oe.param_to_xyz(elev[1], elev[2], elev[3])
            else:
                tX, tY, tZ = elev[1], elev[2], elev[3]
            lb.elevationX[good] = tX
            lb.elevationY[good] = tY
            lb.elevationZ[good] = tZ
    else:
        tMin[good], tMax[good] = _bracketing(
            oe, lb.x[good], lb.y[good], lb.z[good], lb.a[good], lb.b[good], lb.c[good], surfPhysX, surfPhysY,
            mainPartForBracketing, info, isMulti=isMulti, local_n=local_n, invertNormal=invertNormal)
    if noIntersectionSearch:  # reflect.py:676-681
        tMax[good] = 0.
        if oe.isParametric:
            lb.x[good], lb.y[good], lb.z[good] = oe.xyz_to_param(lb.x[good], lb.y[good], lb.z[good])
        _lost = np.zeros(int(good.sum()), dtype=bool)
    else:
        own = getattr(oe.surfParams.get('obj', None), 'find_intersection', None) if oe.surface == 'custom' else None
        if own is not None:  # an element with its own intersection method (BlazedGrating, gratings.py:489-520)
            res_find = own(tMin[good], tMax[good], lb.x[good], lb.y[good], lb.z[good], lb.a[good], lb.b[good],
                           lb.c[good], invertNormal)
        else:
            res_find = find_intersection(oe, None, tMin[good], tMax[good], lb.x[good], lb.y[good], lb.z[good],
                                         lb.a[good], lb.b[good], lb.c[good], invertNormal, info)
        tMax[good], lb.x[good], lb.y[good], lb.z[good] = res_find[:4]
        _lost = res_find[4] if len(res_find) > 4 else np.zeros(int(good.sum()), dtype=bool)
    if oe.isParametric:
        tX, tY, tZ = oe.param_to_xyz(lb.x[good], lb.y[good], lb.z[good])
    else:
        tX, tY, tZ = lb.x[good], lb.y[good], lb.z[good]
    gNormalFZP = None
    lb.state[good] = rays_good(oe, tX, tY, surfPhysX, surfPhysY, surfOptX, surfOptY)
    fzp = getattr(oe, 'fzp', None)
    if fzp is not None:  # use_rays_good_gn (reflect.py:697-698; NormalFZP.rays_good_gn, gratings.py:120-137)
        lb.state[good], gNormalFZP = fzp.rays_good_gn(lb.state[good], tX, tY, oe.lostNum)
    lb.state[np.where(good)[0][_lost]] = oe.lostNum
    goodN = (lb.state == 1)
    goodNsum = goodN.sum()
    if goodNsum > 0:
        lb.path[goodN] += tMax[goodN]
        matSur = material
        isCrystal = material is not None and material.kind == 'crystal'
        oeNormal = list(local_n(lb.x[goodN], lb.y[goodN]))
        fe = getattr(oe, 'figureError', None)
        if fe is not None:  # reflect.py:767-776: d_pitch, d_roll rotations of the surface normal
            oeNormal = [np.asarray(v) * np.ones(int(goodNsum)) for v in oeNormal]
            if oe.isParametric:  # base.py:765-768: the slopes are taken at the Cartesian image of (s, phi)
                rr = oe.local_r(lb.x[goodN], lb.y[goodN])
                xx, yy, _ = oe.param_to_xyz(lb.x[goodN], lb.y[goodN], rr)
                n_distorted = fe.local_n_distorted(xx, yy)
            else:
                n_distorted = fe.local_n_distorted(lb.x[goodN], lb.y[goodN])
            cosX, sinX = np.cos(n_distorted[0]), np.sin(n_distorted[0])
            oeNormal[-2], oeNormal[-1] = ob.rotate_x(oeNormal[-2], oeNormal[-1], cosX, sinX)
            cosY, sinY = np.cos(n_distorted[1]), np.sin(n_distorted[1])
            oeNormal[-3], oeNormal[-1] = ob.rotate_y(oeNormal[-3], oeNormal[-1], cosY, sinY)
        isAsymmetric = len(oeNormal) == 6
        oeNormal = np.asarray(oeNormal, order='F')
        beamInDotNormal = lb.a[goodN]*oeNormal[0] + lb.b[goodN]*oeNormal[1] + lb.c[goodN]*oeNormal[2]
        lb.theta = np.zeros_like(lb.x)
        beamInDotNormal[beamInDotNormal < -1] = -1
        beamInDotNormal[beamInDotNormal > 1] = 1
        lb.theta[goodN] = np.arccos(beamInDotNormal) - np.pi/2
        if isAsymmetric:
            beamInDotSurfaceNormal = lb.a[goodN]*oeNormal[-3] + lb.b[goodN]*oeNormal[-2] + lb.c[goodN]*oeNormal[-1]
        else:
            beamInDotSurfaceNormal = beamInDotNormal
        isMultilayer = material is not None and material.kind == 'multilayer'
        transmitted = (isCrystal or isMultilayer) and matSur.geom.endswith('transmitted')  # toWhere = 2 (reflect.py:749-757)
        isGrating = material is not None and material.kind == 'grating'  # toWhere = 3 (reflect.py:743-744)
        isFZP = material is not None and material.kind == 'FZP'  # toWhere = 4 (reflect.py:745-746)
        isRefr = material is not None and material.kind in ('plate', 'lens')  # toWhere = 1 (reflect.py:732-733)
        if isRefr:
            # reflect.py:894-919
            refractive_index = matSur.get_refractive_index(lb.E[goodN]).real
            if fromVacuum:
                n1overn2 = 1. / refractive_index
            else:
                n1overn2 = refractive_index
            signN = np.sign(-beamInDotNormal)
            n1overn2cosTheta1 = -n1overn2 * beamInDotNormal
            cosTheta2 = signN * np.sqrt(1 - n1overn2**2 + n1overn2cosTheta1**2)
            dn = (n1overn2cosTheta1 - cosTheta2)
            a_out = lb.a[goodN] * n1overn2 + oeNormal[0]*dn
            b_out = lb.b[goodN] * n1overn2 + oeNormal[1]*dn
            c_out = lb.c[goodN] * n1overn2 + oeNormal[2]*dn
        elif isFZP:
            # reflect.py:840-861 with the g vectors rays_good_gn returned (they belong to the rays it kept: state 1)
            a_out, b_out, c_out = _grating_deflection(goodN, lb, np.asarray(gNormalFZP, order='F'), oeNormal,
                                                      beamInDotSurfaceNormal, oe.order, 1)
        elif isGrating:
            # reflect.py:840-861
            if oe.isParametric:
                tXN, tYN = oe.param_to_xyz(lb.x[goodN], lb.y[goodN], lb.z[goodN])[0:2]
            else:
                tXN, tYN = lb.x[goodN], lb.y[goodN]
            gNormal = np.asarray(oe.local_g(tXN, tYN), order='F')
            a_out, b_out, c_out = _grating_deflection(goodN, lb, gNormal, oeNormal, beamInDotSurfaceNormal,
                                                      oe.order, -1)
        elif (isCrystal or isMultilayer) and not transmitted:
            # _get_asymmetric_reflection_grating (reflect.py:568-612)
            normalDotSurfNormal = oeNormal[0]*oeNormal[-3] + oeNormal[1]*oeNormal[-2] + oeNormal[2]*oeNormal[-1]
            bdn = beamInDotNormal.sum() / len(beamInDotNormal)
            sgbdn = 1 if bdn < 0 else -1
            wH = 0
            crystd = matSur.d
            wHd = (1 + wH) / (crystd * 1e-7)
            gNormalCryst = np.asarray((
                (oeNormal[0]-normalDotSurfNormal*oeNormal[-3]) * wHd,
                (oeNormal[1]-normalDotSurfNormal*oeNormal[-2]) * wHd,
                (oeNormal[2]-normalDotSurfNormal*oeNormal[-1]) * wHd), order='F') * sgbdn
            sg = 1 if matSur.geom.startswith('Laue') else -1  # reflect.py:608
            a_out, b_out, c_out = _grating_deflection(goodN, lb, gNormalCryst, oeNormal, beamInDotSurfaceNormal,
                                                      1, sg)
        else:
            a_out = lb.a[goodN] - oeNormal[0]*2*beamInDotNormal
            b_out = lb.b[goodN] - oeNormal[1]*2*beamInDotNormal
            c_out = lb.c[goodN] - oeNormal[2]*2*beamInDotNormal
        if not transmitted:  # toWhere = 2: the ray passes straight
            lb.a[goodN] = a_out
            lb.b[goodN] = b_out
            lb.c[goodN] = c_out
        rollAngle = roll + np.arctan2(oeNormal[-3], oeNormal[-1])
        localJ = ob.rotate_coherency_matrix(lb, goodN, -rollAngle)
        if hasattr(lb, 'Es'):
            cosY, sinY = np.cos(rollAngle), np.sin(rollAngle)
            lb.Es[goodN], lb.Ep[goodN] = ob.rotate_y(lb.Es[goodN], lb.Ep[goodN], cosY, -sinY)
        if material is not None:
            if isCrystal:
                beamOutDotSurfaceNormal = a_out*oeNormal[-3] + b_out*oeNormal[-2] + c_out*oeNormal[-1]
                refl = matSur.get_amplitude(lb.E[goodN], beamInDotSurfaceNormal, beamOutDotSurfaceNormal,
                                            beamInDotNormal)
            elif isMultilayer:  # reflect.py:999-1003
                refl = matSur.get_amplitude(lb.E[goodN], beamInDotSurfaceNormal, lb.x[goodN], lb.y[goodN])
            elif not hasattr(matSur, 'get_amplitude') or isFZP:
                # EmptyMaterial: findReflectivity stays False (:940-942); kind 'FZP': (1, 1, 0) (material.py:457-458)
                refl = 1., 1.
            elif isGrating and matSur.efficiency is not None:  # reflect.py:1005-1011
                refl = matSur.get_grating_efficiency(lb, goodN)
            else:
                refl = matSur.get_amplitude(lb.E[goodN], beamInDotNormal, fromVacuum)
        else:
            refl = 1., 1.
        ras, rap = refl[0], refl[1]
        if isinstance(ras, np.ndarray):
            ras[np.isnan(ras)] = 0.
            rap[np.isnan(rap)] = 0.
        lb.Jss[goodN] = (localJ[0] * ras * np.conjugate(ras)).real
        lb.Jpp[goodN] = (localJ[1] * rap * np.conjugate(rap)).real
        lb.Jsp[goodN] = localJ[2] * ras * np.conjugate(rap)
        if hasattr(lb, 'Es'):
            lb.Es[goodN] *= ras
            lb.Ep[goodN] *= rap
        if (not fromVacuum) and material is not None and material.kind not in ('crystal', 'multilayer'):
            # reflect.py:1048-1059: tMax in mm, refl[2] = mu0 in 1/cm, refl[3] = n.real * k in 1/cm
            att = np.exp(-refl[2] * tMax[goodN] * 0.1)
            lb.Jss[goodN] *= att
            lb.Jpp[goodN] *= att
            lb.Jsp[goodN] *= att
            if hasattr(lb, 'Es'):
                mPh = att**0.5 * np.exp(0.1j * refl[3] * tMax[goodN])
                lb.Es[goodN] *= mPh
                lb.Ep[goodN] *= mPh
        elif hasattr(lb, 'Es'):
            mPh = np.exp(1e7j * lb.E[goodN]/CHBAR * tMax[goodN])
            lb.Es[goodN] *= mPh
            lb.Ep[goodN] *= mPh
    if oe.isParametric:
        lb.s = np.copy(lb.x)
        lb.phi = np.copy(lb.y)
        lb.r = np.copy(lb.z)
        lb.x[good], lb.y[good], lb.z[good] = oe.param_to_xyz(lb.x[good], lb.y[good], lb.z[good])
    if goodNsum > 0:
        vlb.Jss[goodN], vlb.Jpp[goodN], vlb.Jsp[goodN] = ob.rotate_coherency_matrix(lb, goodN, rollAngle)
        if hasattr(lb, 'Es'):
            vlb.Es[goodN], vlb.Ep[goodN] = ob.rotate_y(lb.Es[goodN], lb.Ep[goodN], cosY, sinY)
    if vlb is not lb:
        ob.copy_beam(vlb, lb, good, includeState=True, includeJspEsp=False)
    if dx:
        vlb.x[good] += dx
    if dy:
        vlb.y[good] += dy
    if dz:
        vlb.z[good] += dz
    if oe.extraPitch or oe.extraRoll or oe.extraYaw:
        ob.rotate_beam(vlb, good, rotationSequence='-'+oe.extraRotationSequence,
                       pitch=extraAnglesSign*oe.extraPitch, roll=oe.extraRoll, yaw=extraAnglesSign*oe.extraYaw)
    ob.rotate_beam(vlb, good, rotationSequence='-'+oe.rotationSequence, pitch=pitch, roll=roll, yaw=yaw)
    if is2ndXtal:
        ob.rotate_beam(vlb, good, roll=np.pi)
    if info is not None:
        info['footprint'] = np.hstack((
            np.min(np.vstack((vlb.x[good], vlb.y[good], vlb.z[good])), axis=1),
            np.max(np.vstack((vlb.x[good], vlb.y[good], vlb.z[good])), axis=1))).reshape(2, 3)


def reflect(oe, beam, needLocal=True, createdByDiffract=False, info=None, noIntersectionSearch=False):
    """OE.reflect (reflect.py:104-135). Returns gb, lb."""
    gb = ob.Beam(copyFrom=beam)
    lb = ob.Beam(copyFrom=beam) if needLocal else gb
    good = beam.state > 0
    if good.sum() == 0:
        return gb, lb
    ob.global_to_virgin_local(oe.azimuth, beam, lb, oe.center, good)
    reflect_local(oe, good, lb, gb, oe.pitch, oe.roll, oe.yaw, oe.dx, material=oe.material, info=info,
                  noIntersectionSearch=noIntersectionSearch)
    if createdByDiffract:
        goodAfter = gb.state == 1
    else:
        goodAfter = (gb.state == 1) | (gb.state == 2)
    if goodAfter.sum() > 0:
        ob.virgin_local_to_global(oe.azimuth, gb, oe.center, goodAfter)
    notGood = ~goodAfter
    if notGood.sum() > 0:
        ob.copy_beam(gb, beam, notGood)
    return gb, lb


def multiple_reflect(oe, beam, maxReflections=1000, infos=None, needElevationMap=False):
    """OE.multiple_reflect (reflect.py:165-264).  gb and lb are ONE object there, so every pass works in place on the
    virgin-local beam; lbN stacks the passes with Beam.concatenate (beams.py:230-295).  needElevationMap: elevationD /
    X / Y / Z of the passes after the first (reflect.py:212-216, :651-659)."""
    gb = ob.Beam(copyFrom=beam)
    lb = gb
    good = beam.state > 0
    if good.sum() == 0:
        return gb, lb
    ob.global_to_virgin_local(oe.azimuth, beam, lb, oe.center, good)
    iRefl = 0
    isMulti = False
    while iRefl < maxReflections:
        tmpX, tmpY, tmpZ = np.copy(lb.x[good]), np.copy(lb.y[good]), np.copy(lb.z[good])
        info = {} if infos is not None else None
        if iRefl == 0 and needElevationMap:
            lb.elevationD = -np.ones_like(lb.x)
            lb.elevationX = -np.ones_like(lb.x)*maxHalfSizeOfOE
            lb.elevationY = -np.ones_like(lb.x)*maxHalfSizeOfOE
            lb.elevationZ = -np.ones_like(lb.x)*maxHalfSizeOfOE
        reflect_local(oe, good, lb, gb, oe.pitch, oe.roll, oe.yaw, oe.dx, material=oe.material, info=info,
                      isMulti=isMulti, needElevationMap=needElevationMap)
        if infos is not None:
            infos.append(info)
        if iRefl == 0:
            isMulti = True
            lb.nRefl = np.zeros_like(lb.state)
        ov = lb.state[good] == 3
        lb.x[np.where(good)[0][ov]] = tmpX[ov]
        lb.y[np.where(good)[0][ov]] = tmpY[ov]
        lb.z[np.where(good)[0][ov]] = tmpZ[ov]
        good = (lb.state == 1) | (lb.state == 2)
        lb.nRefl[good] += 1
        if iRefl == 0:
            lbN = ob.Beam(copyFrom=lb)
        else:
            ob.concatenate(lbN, lb)
        iRefl += 1
        if good.sum() == 0:
            break
    goodAfter = gb.nRefl > 0
    gb.state[goodAfter] = 1
    if goodAfter.sum() > 0:
        ob.virgin_local_to_global(oe.azimuth, gb, oe.center, goodAfter)
    notGood = ~goodAfter
    if notGood.sum() > 0:
        ob.copy_beam(gb, beam, notGood)
    return gb, lbN


def double_reflect(dcm, beam, needLocal=True, info1=None, info2=None, fromVacuum1=True, fromVacuum2=True):
    """DCM.double_reflect (dcm.py:273-354). `dcm` is an OESpec with the extra DCM fields:
    bragg, cryst1roll, cryst2roll, cryst2pitch, cryst2finePitch, cryst2perpTransl, cryst2longTransl, material2,
    surfPhysX2, surfPhysY2, surfOptX2, surfOptY2.  Returns gb2, lo1, lo2."""
    gb = ob.Beam(copyFrom=beam)
    lo1 = ob.Beam(copyFrom=beam) if needLocal else gb
    good1 = beam.state > 0
    if good1.sum() == 0:
        return gb, lo1, lo1
    ob.global_to_virgin_local(dcm.azimuth, beam, lo1, dcm.center, good1)
    reflect_local(dcm, good1, lo1, gb, dcm.pitch + dcm.bragg, dcm.roll + dcm.cryst1roll, dcm.yaw, dcm.dx,
                  local_n=dcm.local_n, material=dcm.material, info=info1, fromVacuum=fromVacuum1)
    goodAfter1 = (gb.state == 1) | (gb.state == 2)
    notGood = ~goodAfter1
    if notGood.sum() > 0:
        ob.copy_beam(gb, beam, notGood)
    gb2 = ob.Beam(copyFrom=gb)
    lo2 = ob.Beam(copyFrom=gb2) if needLocal else gb2
    good2 = goodAfter1
    if needLocal and (~good2).sum() > 0:
        lo2.state[~good2] = 0
        lo2.x[~good2] = 0.
        lo2.y[~good2] = 0.
        lo2.z[~good2] = 0.
    if hasattr(dcm, 't'):  # is instance of Plate (dcm.py:311-312)
        gb2.state[~good2] = dcm.lostNum
    if good2.sum() == 0:
        return gb2, lo1, lo2

    def local_n2(x, y):
        res = dcm.local_n(x, y)
        if dcm.alpha:
            res[1] *= -1
        return res

    reflect_local(dcm, good2, lo2, gb2,
                  -dcm.pitch - dcm.bragg + dcm.cryst2pitch + dcm.cryst2finePitch,
                  dcm.roll + dcm.cryst2roll, -dcm.yaw, -dcm.dx, dcm.cryst2longTransl, -dcm.cryst2perpTransl,
                  local_n=local_n2, material=dcm.material2, is2ndXtal=True,
                  surfPhysX=dcm.surfPhysX2, surfPhysY=dcm.surfPhysY2, surfOptX=dcm.surfOptX2,
                  surfOptY=dcm.surfOptY2, info=info2, fromVacuum=fromVacuum2)
    goodAfter2 = (gb2.state == 1) | (gb2.state == 2)
    ob.virgin_local_to_global(dcm.azimuth, gb2, dcm.center, goodAfter2)
    notGood = ~goodAfter2
    if hasattr(dcm, 't'):  # dcm.py:331-332
        gb2.state[notGood] = dcm.lostNum
    if notGood.sum() > 0:
        ob.copy_beam(gb2, beam, notGood)
    return gb2, lo1, lo2


def double_refract(plate, beam, needLocal=True):
    """Plate.double_refract (refractive.py:171-234).  `plate` is a DCM-like OESpec with t (cryst2perpTransl = -t,
    cryst2pitch = wedgeAngle, limits 2 mirrored in x) and a material of kind 'plate' / 'lens'."""
    return double_reflect(plate, beam, needLocal, fromVacuum1=True, fromVacuum2=False)


def rotate_point(point, rotationSequence='RzRyRx', pitch=0, roll=0, yaw=0):
    """raycing.rotate_point (_rotate.py:87-108)."""
    angles = {'z': yaw, 'y': roll, 'x': pitch}
    rotates = {'z': ob.rotate_z, 'y': ob.rotate_y, 'x': ob.rotate_x}
    ind1 = {'z': 0, 'y': 0, 'x': 1}
    ind2 = {'z': 1, 'y': 2, 'x': 2}
    newp = [coord for coord in point]
    if rotationSequence[0] == '-':
        seq = rotationSequence[6] + rotationSequence[4] + rotationSequence[2]
    else:
        seq = rotationSequence[1] + rotationSequence[3] + rotationSequence[5]
    for sq in seq:
        angle, rotate = angles[sq], rotates[sq]
        if angle != 0:
            cA = np.cos(angle)
            sA = np.sin(angle)
            newp[ind1[sq]], newp[ind2[sq]] = rotate(newp[ind1[sq]], newp[ind2[sq]], cA, sA)
    return newp


def multiple_refract(lens, beam, needLocal=True):
    """ParaboloidFlatLens.multiple_refract (refractive.py:457-549); lens.nCRL, lens.zmax, lens.isDouble."""
    if lens.nCRL == 1:
        return double_refract(lens, beam, needLocal)
    tempCenter = [c for c in lens.center]
    beamIn = beam
    zmax = 5 if lens.zmax is None else lens.zmax
    step = 2.*zmax + lens.t if lens.isDouble else zmax + lens.t
    for ilens in range(lens.nCRL):
        lglobal, tlocal1, tlocal2 = double_refract(lens, beamIn, needLocal)
        if lens.zmax is not None:
            toward = rotate_point([0, 0, 1], lens.rotationSequence, lens.pitch, lens.roll, lens.yaw)
            lens.center[0] -= step * toward[0]
            lens.center[1] -= step * toward[1]
            lens.center[2] -= step * toward[2]
        beamIn = lglobal
        if ilens == 0:
            llocal1, llocal2 = tlocal1, tlocal2
    lens.center = tempCenter
    return lglobal, llocal1, llocal2
