# -*- coding: utf-8 -*-
"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): numpy restatements of the reference's curved optical elements
whose surface the product compiles from spliced CUDA texts (xrt_b200/oes_spliced.py).  Each class gives
local_z(x, y) and local_n(x, y) in the reference's operation order; oracle.reflect.OESpec(surface='custom',
surfParams={'obj': <instance>}) plugs them into the common intersection / reflection machinery.

Pinned bit for bit on the reference's own outputs (tests/golden/reflect_spliced_*.npz, tools/gen_golden.py
section `spliced`)."""
import numpy as np

from . import beam as ob


class BentFlatMirror(object):
    """oes/__init__.py:240-304: meridionally bent parabolic cylinder with fixed ends."""

    def __init__(self, R, limPhysY0):
        self.R = R
        self.limPhysY0 = limPhysY0

    def local_z(self, x, y):  # :289-293
        return (y**2 - self.limPhysY0**2) / 2.0 / self.R

    def local_n(self, x, y):  # :295-304
        a = 0.
        b = -y / self.R
        c = 1.
        norm = (b**2 + 1)**0.5
        return [a/norm, b/norm, c/norm]


class ConicalMirror(object):
    """oes/__init__.py:589-636."""

    def __init__(self, L0, theta):
        self.L0 = L0
        self.tt = np.tan(theta)  # :610-615
        self.t2t = np.tan(2*theta)
        self.redfocus = np.cos(theta)**2 / (1./self.tt - 1./self.t2t)

    def local_z(self, x, y):  # :623-627
        sqroot = np.sqrt(0.25*self.t2t**2*(y - self.L0)**2 - self.redfocus*self.t2t*x**2)
        z = -0.5*self.t2t*(y-self.L0) - np.sign(self.t2t)*sqroot
        return z

    def local_n(self, x, y):  # :629-636
        sqroot = np.sign(self.t2t)*np.sqrt(0.25*self.t2t**2*(y - self.L0)**2 - self.redfocus*x*x*self.t2t)
        a = -x*self.redfocus*self.t2t/sqroot
        b = .5*self.t2t + 0.25*self.t2t**2*(y-self.L0)/sqroot
        c = 1.
        norm = (a**2 + b**2 + 1.)**0.5
        return [a/norm, b/norm, c/norm]


class JohannCylinder(object):
    """oes/bragg.py:104-176: simply bent reflective crystal."""

    def __init__(self, Rm, crossSection='circular', alpha=None):
        self.Rm = Rm
        self.crossSection = crossSection
        self.alpha = alpha
        if alpha:
            self.cosalpha, self.sinalpha = np.cos(alpha), np.sin(alpha)

    def local_z(self, x, y):  # :145-152
        if self.crossSection.startswith('circ'):
            sq = self.Rm**2 - y**2
            return self.Rm - np.sqrt(sq)
        else:
            return y**2 / 2.0 / self.Rm

    def local_n_cylinder(self, x, y, R, alpha):  # :154-171
        a = np.zeros_like(x)
        b = -y / R
        if self.crossSection.startswith('circ'):
            c = (R**2 - y**2)**0.5 / R
        else:
            norm = (b**2 + 1)**0.5
            b /= norm
            c = 1. / norm
        if alpha:
            bAlpha, cAlpha = ob.rotate_x(b, c, self.cosalpha, -self.sinalpha)
            return [a, bAlpha, cAlpha, a, b, c]
        else:
            return [a, b, c]

    def local_n(self, x, y):  # :173-176
        return self.local_n_cylinder(x, y, self.Rm, self.alpha)


class JohanssonCylinder(JohannCylinder):
    """oes/bragg.py:179-197: ground-bent crystal, exact expressions."""

    def local_n(self, x, y):
        nSurf = self.local_n_cylinder(x, y, self.Rm, None)
        a = np.zeros_like(x)
        b = -y
        c = (self.Rm**2 - y**2)**0.5 + self.Rm
        if self.alpha:
            b, c = ob.rotate_x(b, c, self.cosalpha, -self.sinalpha)
        norm = np.sqrt(b**2 + c**2)
        return [a/norm, b/norm, c/norm, nSurf[-3], nSurf[-2], nSurf[-1]]


def _rotate_y(x, z, cosangle, sinangle):  # _rotate.py:11-14
    return cosangle*x + sinangle*z, -sinangle*x + cosangle*z


class JohannToroid(object):
    """oes/bragg.py:200-269: 2D bent reflective crystal."""

    def __init__(self, Rm, Rs=None, alpha=None):
        self.Rm = Rm
        self.Rs = Rs if Rs is not None else Rm
        self.alpha = alpha
        if alpha:
            self.cosalpha, self.sinalpha = np.cos(alpha), np.sin(alpha)

    def local_z(self, x, y):  # :238-243
        z = self.Rm - self.Rs - (self.Rm**2 - y**2)**0.5
        cosangle, sinangle = (z**2 - x**2)**0.5 / abs(z), -x/abs(z)
        bla, z = _rotate_y(0, z, cosangle, sinangle)
        return z + self.Rs

    def local_n(self, x, y):  # :245-269
        Rm, Rs, alpha = self.Rm, self.Rs, self.alpha
        a = np.zeros_like(x)
        b = -y / Rm
        c = (Rm**2 - y**2)**0.5 / Rm
        if alpha:
            aAlpha = np.zeros_like(x)
            bAlpha, cAlpha = ob.rotate_x(b, c, self.cosalpha, -self.sinalpha)
        r = Rs - (Rm - (Rm**2 - y**2)**0.5)
        cosangle, sinangle = (r**2 - x**2)**0.5 / r, -x/r
        a, c = _rotate_y(a, c, cosangle, sinangle)
        if alpha:
            aAlpha, cAlpha = _rotate_y(aAlpha, cAlpha, cosangle, sinangle)
            return [aAlpha, bAlpha, cAlpha, a, b, c]
        else:
            return [a, b, c]


class ParabolicalMirrorParam(object):
    """oes/parametric.py:252-476: paraboloid / parabolic cylinder in the parametric (s, phi, r) representation; the
    scalars parabParam, y0, z0, cosGamma, sinGamma are those of _reset_pq (:405-430)."""

    def __init__(self, parabParam, y0, z0, cosGamma, sinGamma, isCylindrical=False):
        self.parabParam, self.y0, self.z0 = parabParam, y0, z0
        self.cosGamma, self.sinGamma = cosGamma, sinGamma
        self.isCylindrical = isCylindrical

    def xyz_to_param(self, x, y, z):  # :441-444
        yNew, zNew = ob.rotate_x(y - self.y0, z - self.z0, self.cosGamma, self.sinGamma)
        return yNew, np.arctan2(x, zNew), np.sqrt(x**2 + zNew**2)

    def param_to_xyz(self, s, phi, r):  # :446-451
        x = r * np.sin(phi)
        y = s
        z = r * np.cos(phi)
        yNew, zNew = ob.rotate_x(y, z, self.cosGamma, -self.sinGamma)
        return x, yNew + self.y0, zNew + self.z0

    def local_r(self, s, phi):  # :453-461 (isClosed = False)
        r2 = self.parabParam*s + self.parabParam**2
        r2[r2 < 0] = 0
        r = 2 * r2**0.5
        if self.isCylindrical:
            r /= abs(np.cos(phi))
        return np.where(abs(phi) > np.pi/2, r, np.ones_like(phi)*1e20)

    def local_n(self, s, phi):  # :463-474
        nr = self.parabParam / (self.parabParam*s + self.parabParam**2)**0.5
        norm = np.sqrt(nr**2 + 1)
        b = nr / norm
        if self.isCylindrical:
            a = np.zeros_like(phi)
            c = 1. / norm
        else:
            a = -np.sin(phi) / norm
            c = -np.cos(phi) / norm
        bNew, cNew = ob.rotate_x(b, c, self.cosGamma, -self.sinGamma)
        return [a, bNew, cNew]


class FigureErrorSpline(object):
    """FigureErrorBase.build_spline / local_z_distorted / local_n_distorted (figure_error.py:207-265): a height map in
    nm on (y1d, x1d) -> scipy RectBivariateSpline (kx = ky = 3, the primitive the reference itself uses); heights in
    mm, slopes as (d_pitch, d_roll)."""

    def __init__(self, x1d, y1d, z_nm, xShift=0., yShift=0., splineOrder=3):
        from scipy import interpolate
        self.local_z_spline = interpolate.RectBivariateSpline(y1d, x1d, z_nm, kx=splineOrder, ky=splineOrder)
        self.xShift, self.yShift = xShift, yShift

    def local_z_distorted(self, x, y):
        return self.local_z_spline.ev(y+self.yShift, x+self.xShift) * 1e-6

    def local_n_distorted(self, x, y):
        a = self.local_z_spline.ev(y+self.yShift, x+self.xShift, dx=0, dy=1) * 1e-6
        b = self.local_z_spline.ev(y+self.yShift, x+self.xShift, dx=1, dy=0) * 1e-6
        return [np.arctan(b), -np.arctan(a)]


class LauePlate(object):
    """oes/laue.py:11-24: a flat Laue plate -- the normal to the diffracting planes lies (nearly) in the surface."""

    def __init__(self, alpha=None):
        self.alpha = alpha
        if alpha:
            self.cosalpha, self.sinalpha = np.cos(alpha), np.sin(alpha)

    def local_z(self, x, y):
        return np.zeros_like(y)

    def local_n(self, x, y):
        a, b, c = 0, 0, 1
        if self.alpha:
            bB, cB = ob.rotate_x(b, c, -self.sinalpha, -self.cosalpha)
        else:
            bB, cB = c, -b
        return [a, bB, cB, a, b, c]


class BentLaueCylinder(object):
    """oes/laue.py:26-175 (local_z :143-148, local_n_cylinder :150-169): cylindrically bent Laue crystal."""

    def __init__(self, R, crossSection='parabolic', alpha=None):
        self.R, self.crossSection, self.alpha = R, crossSection, alpha
        if alpha:
            self.cosalpha, self.sinalpha = np.cos(alpha), np.sin(alpha)

    def local_z(self, x, y):
        if self.crossSection.startswith('circ'):
            return self.R - np.sqrt(self.R**2 - y**2)
        else:
            return y**2 / 2.0 / self.R

    def local_n(self, x, y):
        R = self.R
        a = np.zeros_like(x)
        b = -y / R
        if self.crossSection.startswith('circ'):
            c = (R**2 - y**2)**0.5 / R
        else:
            norm = (b**2 + 1)**0.5
            b /= norm
            c = 1. / norm
        if self.alpha:
            bB, cB = ob.rotate_x(b, c, -self.sinalpha, -self.cosalpha)
        else:
            bB, cB = c, -b
        return [a, bB, cB, a, b, c]


class NormalFZP(object):
    """NormalFZP (oes/gratings.py:10-137): zone radii rn (reset, :94-117) and rays_good_gn (:120-137): rays in opaque
    zones are lost, the others get the local radial groove vector."""

    def __init__(self, f, E, N, isCentralZoneBlack=True):
        from scipy import interpolate
        from .physconsts import CH
        self.isCentralZoneBlack = isCentralZoneBlack
        lambdaE = CH / E * 1e-7
        self.zones = np.arange(N+1)
        self.rn = np.sqrt(self.zones*f*lambdaE + 0.25*(self.zones*lambdaE)**2)
        self.r_to_i = interpolate.interp1d(self.rn, self.zones, bounds_error=False, fill_value=0)
        self.i_to_r = interpolate.interp1d(self.zones, self.rn, bounds_error=False, fill_value=0)

    def rays_good_gn(self, locState, x, y, lostNum):
        r = np.sqrt(x**2 + y**2)
        i = (self.r_to_i(r)).astype(int)
        good = ((i % 2 == int(self.isCentralZoneBlack)) & (r < self.rn[-1]) & (locState == 1))
        locState[~good] = lostNum
        gz = np.zeros_like(x[good])
        rho = 1./(self.i_to_r(i[good]+1) - self.i_to_r(i[good]-1))
        gx = -x[good] / r[good] * rho
        gy = -y[good] / r[good] * rho
        return locState, (gx, gy, gz)


class BlazedGrating(object):
    """BlazedGrating with a constant line density (oes/gratings.py:316-535: reset :416-441, local_pre :459-473,
    local_z :475-480, local_n :482-490, find_intersection :492-520)."""

    def __init__(self, blaze, antiblaze, rho0):
        self.blaze, self.antiblaze, self.rho0 = blaze, antiblaze, rho0
        self.rho_1 = 1. / rho0
        self.sinBlaze, self.cosBlaze, self.tanBlaze = np.sin(blaze), np.cos(blaze), np.tan(blaze)
        self.sinAntiblaze, self.cosAntiblaze, self.tanAntiblaze = \
            np.sin(antiblaze), np.cos(antiblaze), np.tan(antiblaze)

    def local_pre(self, x, y):
        y0 = (y // self.rho_1) * self.rho_1
        y1 = y0 + self.rho_1
        yL = y % self.rho_1
        yC = (y1-y0) / (1 + self.tanAntiblaze/self.tanBlaze)
        return 0, y0, y1, yC, yL

    def local_z(self, x, y):
        y0ind, y0, y1, yC, yL = self.local_pre(x, y)
        return np.where(yL > yC, -(y1-y) * self.tanBlaze, -yL * self.tanAntiblaze)

    def local_n(self, x, y):
        y0ind, y0, y1, yC, yL = self.local_pre(x, y)
        return [np.zeros_like(x), np.where(yL > yC, -self.sinBlaze, self.sinAntiblaze),
                np.where(yL > yC, self.cosBlaze, self.cosAntiblaze)]

    def find_intersection(self, t1, t2, x, y, z, a, b, c, invertNormal):
        b_c = b / c
        n = np.floor((y - b_c*z) / self.rho_1)
        y0 = self.rho_1 * n
        y1 = y0 + self.rho_1
        if self.antiblaze == np.pi/2:
            zabl = (y0-y) / b_c + z
        else:
            zabl = -self.tanAntiblaze * (y - b_c*z - y0) / (1 + self.tanAntiblaze*b_c)
        if self.blaze == np.pi/2:
            zbl = (y1-y) / b_c + z
        else:
            zbl = self.tanBlaze * (y - b_c*z - y1) / (1 - self.tanBlaze*b_c)
        zabl[zabl > 0] = zbl[zabl > 0] - 1
        zbl[zbl > 0] = zabl[zbl > 0] - 1
        z2 = zbl
        y2 = b_c * (z2 - z) + y
        t2 = (y2 - y) / b
        x2 = x + t2 * a
        return t2, x2, y2, z2
