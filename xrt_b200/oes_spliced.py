# -*- coding: utf-8 -*-
"""Curved optical elements of xrt whose surfaces are spliced CUDA texts (xrt_b200/custom.py), with xrt's class names
and constructor arguments:

    BentFlatMirror(..., R) = SimpleVCM                      oes/__init__.py:240-306
    ConicalMirror(..., L0, theta)                           oes/__init__.py:589-636
    JohannCylinder(..., Rm, crossSection)                   oes/bragg.py:104-176
    JohanssonCylinder(..., Rm, crossSection)                oes/bragg.py:179-197
    JohannToroid(..., Rm, Rs)                               oes/bragg.py:200-269
    ParabolicalMirrorParam(..., p | q, isCylindrical)       oes/parametric.py:252-476  (parametric: s, phi, r)
    BentLaueCylinder(..., R, crossSection)                  oes/laue.py:26-175   (planes' normal + surface normal)
    LauePlate(...)                                          oes/laue.py:11-24    (flat: the built-in kernel, no texts)
    BlazedGrating(..., blaze, antiblaze, rho)               oes/gratings.py:316-535  (own analytic find_intersection)

The texts follow the reference's numpy expressions operation by operation (the plug-in is compiled with
-fmad=false); scalars the reference forms from Python floats (`self.Rm**2`, `0.25*self.t2t**2`) are formed on the
host by the same Python expressions and travel in `cu_plist`, so they carry the same roundings.  These classes double
as the worked examples of writing `cu_local_z` / `cu_local_n` for an element of one's own."""
import numpy as np

from .oes import OE, _abs_pitch_to_axis

# rotate_x / rotate_y of _rotate.py:5-14, for the texts below
_ROT = '''
__device__ __forceinline__ void rot_x(double y, double z, double c, double s, double& yo, double& zo)
{ yo = c * y - s * z; zo = s * y + c * z; }
'''


class BentFlatMirror(OE):
    """z = (y^2 - limPhysY[0]^2) / 2R: a meridionally bent parabolic cylinder with fixed ends."""
    cu_plist = ('R', '_y0sq')
    cu_local_z = '''
    __device__ double local_z(const double* cu_plist, double x, double y)
    { return (y * y - cu_plist[1]) / 2.0 / cu_plist[0]; }'''
    cu_local_n = '''
    __device__ void local_n(const double* cu_plist, double x, double y, double* n)
    {
        const double b = -y / cu_plist[0];
        const double norm = sqrt(b * b + 1.0);
        n[0] = 0.0 / norm; n[1] = b / norm; n[2] = 1.0 / norm;
    }'''

    def __init__(self, *args, **kwargs):
        R = kwargs.pop('R', 5.0e6)
        OE.__init__(self, *args, **kwargs)
        if isinstance(R, (list, tuple)):
            raise NotImplementedError('give R as a number (Coddington (p, q) tuples are resolved by xrt)')
        self.R = 1e100 if R is None else R

    @property
    def _y0sq(self):
        return self.limPhysY[0]**2


SimpleVCM = BentFlatMirror


class ConicalMirror(OE):
    """Conical mirror with its base parallel to the side of the cone."""
    cu_plist = ('L0', 't2t', 'redfocus', '_c1', '_c2', '_sgn')
    cu_local_z = '''
    __device__ double local_z(const double* cu_plist, double x, double y)
    {
        const double L0 = cu_plist[0], t2t = cu_plist[1], c1 = cu_plist[3], c2 = cu_plist[4], sgn = cu_plist[5];
        const double yl = y - L0;
        const double sqroot = sqrt(c1 * (yl * yl) - c2 * (x * x));
        return (-0.5 * t2t) * yl - sgn * sqroot;
    }'''
    cu_local_n = '''
    __device__ void local_n(const double* cu_plist, double x, double y, double* n)
    {
        const double L0 = cu_plist[0], t2t = cu_plist[1], rf = cu_plist[2], c1 = cu_plist[3], sgn = cu_plist[5];
        const double yl = y - L0;
        const double sqroot = sgn * sqrt(c1 * (yl * yl) - ((rf * x) * x) * t2t);
        const double a = ((-x) * rf) * t2t / sqroot;
        const double b = 0.5 * t2t + c1 * yl / sqroot;
        const double norm = sqrt(a * a + b * b + 1.0);
        n[0] = a / norm; n[1] = b / norm; n[2] = 1.0 / norm;
    }'''

    def __init__(self, *args, **kwargs):
        self.L0 = kwargs.pop('L0', 1000.)
        theta = kwargs.pop('theta', np.pi/6.)
        OE.__init__(self, *args, **kwargs)
        self.theta = theta

    @property
    def theta(self):
        return self._theta

    @theta.setter
    def theta(self, theta):  # oes/__init__.py:610-615
        self._theta = theta
        self.tt = np.tan(self._theta)
        self.t2t = np.tan(2*self._theta)
        self.redfocus = np.cos(self._theta)**2 / (1./self.tt - 1./self.t2t)

    _c1 = property(lambda self: 0.25*self.t2t**2)
    _c2 = property(lambda self: self.redfocus*self.t2t)
    _sgn = property(lambda self: float(np.sign(self.t2t)))


class JohannCylinder(OE):
    """Simply bent reflective crystal; with alpha the normals (n_B, n_s) are both returned."""
    cu_plist = ('crossSectionInt', 'Rm', '_Rm2', '_cosalpha', '_sinalpha')
    cu_local_z = '''
    __device__ double local_z(const double* cu_plist, double x, double y)
    {
        if (cu_plist[0] == 0.0) return cu_plist[1] - sqrt(cu_plist[2] - y * y);
        return y * y / 2.0 / cu_plist[1];
    }'''
    _cu_cylinder = _ROT + '''
    __device__ __forceinline__ void n_cylinder(const double* cu_plist, double y, double& b, double& c)
    {
        const double R = cu_plist[1];
        b = -y / R;
        if (cu_plist[0] == 0.0) {
            c = sqrt(cu_plist[2] - y * y) / R;
        } else {
            const double norm = sqrt(b * b + 1.0);
            b = b / norm;
            c = 1.0 / norm;
        }
    }'''
    cu_local_n = _cu_cylinder + '''
    __device__ void local_n(const double* cu_plist, double x, double y, double* n)
    {
        double b, c;
        n_cylinder(cu_plist, y, b, c);
        n[0] = 0.0; n[1] = b; n[2] = c;
        if (cu_plist[4] != 0.0 || cu_plist[3] != 1.0) {  // alpha: (n_B, n_s)
            n[3] = 0.0; n[4] = b; n[5] = c;
            rot_x(b, c, cu_plist[3], -cu_plist[4], n[1], n[2]);
        }
    }'''

    def __init__(self, *args, **kwargs):
        self.Rm = kwargs.pop('Rm', 1000.)
        self.crossSection = kwargs.pop('crossSection', 'circular')
        if not (self.crossSection.startswith('circ') or self.crossSection.startswith('parab')):
            raise ValueError('unknown crossSection!')
        self.crossSectionInt = 0 if self.crossSection.startswith('circ') else 1
        OE.__init__(self, *args, **kwargs)

    _Rm2 = property(lambda self: self.Rm**2)
    _cosalpha = property(lambda self: float(np.cos(self.alpha)) if self.alpha else 1.)
    _sinalpha = property(lambda self: float(np.sin(self.alpha)) if self.alpha else 0.)


class JohanssonCylinder(JohannCylinder):
    """Ground-bent (Johansson) crystal: atomic planes on 2 Rm, surface on Rm (exact expressions)."""
    cu_two_normals = True
    cu_local_n = JohannCylinder._cu_cylinder + '''
    __device__ void local_n(const double* cu_plist, double x, double y, double* n)
    {
        double bs, cs;
        n_cylinder(cu_plist, y, bs, cs);
        double b = -y;
        double c = sqrt(cu_plist[2] - y * y) + cu_plist[1];
        if (cu_plist[4] != 0.0 || cu_plist[3] != 1.0) {
            const double b0 = b, c0 = c;
            rot_x(b0, c0, cu_plist[3], -cu_plist[4], b, c);
        }
        const double norm = sqrt(b * b + c * c);
        n[0] = 0.0 / norm; n[1] = b / norm; n[2] = c / norm;
        n[3] = 0.0; n[4] = bs; n[5] = cs;
    }'''


class JohannToroid(OE):
    """2D bent reflective crystal with meridional and sagittal radii Rm, Rs."""
    cu_plist = ('Rm', 'Rs', '_Rm2', '_RmRs', '_cosalpha', '_sinalpha')
    cu_local_z = '''
    __device__ double local_z(const double* cu_plist, double x, double y)
    {
        const double z = cu_plist[3] - sqrt(cu_plist[2] - y * y);
        const double cosangle = sqrt(z * z - x * x) / fabs(z);
        return cosangle * z + cu_plist[1];
    }'''
    cu_local_n = _ROT + '''
    __device__ void local_n(const double* cu_plist, double x, double y, double* n)
    {
        const double Rm = cu_plist[0], Rs = cu_plist[1];
        const double sq = sqrt(cu_plist[2] - y * y);
        const double b = -y / Rm;
        const double c = sq / Rm;
        const double r = Rs - (Rm - sq);
        const double cosangle = sqrt(r * r - x * x) / r, sinangle = -x / r;
        if (cu_plist[5] != 0.0 || cu_plist[4] != 1.0) {
            double bA, cA;
            rot_x(b, c, cu_plist[4], -cu_plist[5], bA, cA);
            n[0] = sinangle * cA; n[1] = bA; n[2] = cosangle * cA;
            n[3] = sinangle * c;  n[4] = b;  n[5] = cosangle * c;
        } else {
            n[0] = sinangle * c; n[1] = b; n[2] = cosangle * c;
        }
    }'''

    def __init__(self, *args, **kwargs):
        self.Rm = kwargs.pop('Rm', 1000.)
        self._Rs = kwargs.pop('Rs', None)
        OE.__init__(self, *args, **kwargs)

    @property
    def Rs(self):
        return self._Rs if self._Rs is not None else self.Rm

    @Rs.setter
    def Rs(self, Rs):
        self._Rs = Rs

    _Rm2 = property(lambda self: self.Rm**2)
    _RmRs = property(lambda self: self.Rm - self.Rs)
    _cosalpha = property(lambda self: float(np.cos(self.alpha)) if self.alpha else 1.)
    _sinalpha = property(lambda self: float(np.sin(self.alpha)) if self.alpha else 0.)


class ParabolicalMirrorParam(OE):
    """Paraboloid / parabolic cylinder in the parametric (s, phi, r) representation: s along the paraboloid axis
    from the focus, (phi, r) polar coordinates normal to it.  One of *p* (collimating) or *q* (focusing); the axis is
    the global y axis (f1 / f2 / parabolaAxis are resolved by xrt)."""
    isParametric = True
    cu_plist = ('parabParam', '_pp2', 'y0', 'z0', 'cosGamma', 'sinGamma', '_isCyl')
    cu_local_z = '''
    __device__ double local_z(const double* cu_plist, double s, double phi)   // = local_r(s, phi)
    {
        double r2 = cu_plist[0] * s + cu_plist[1];
        if (r2 < 0.0) r2 = 0.0;
        double r = 2.0 * sqrt(r2);
        if (cu_plist[6] != 0.0) r /= fabs(cos(phi));
        return (fabs(phi) > 3.141592653589793 / 2) ? r : 1e20;
    }'''
    cu_local_n = _ROT + '''
    __device__ void local_n(const double* cu_plist, double s, double phi, double* n)
    {
        const double nr = cu_plist[0] / sqrt(cu_plist[0] * s + cu_plist[1]);
        const double norm = sqrt(nr * nr + 1.0);
        const double b = nr / norm;
        double a, c;
        if (cu_plist[6] != 0.0) { a = 0.0; c = 1.0 / norm; }
        else { a = -sin(phi) / norm; c = -cos(phi) / norm; }
        n[0] = a;
        rot_x(b, c, cu_plist[4], -cu_plist[5], n[1], n[2]);
    }'''
    cu_xyz_param = '''  // (rot_x comes with cu_local_n above: the texts share one header)
    __device__ void xyz_to_param(const double* cu_plist, double& x, double& y, double& z)
    {
        double yN, zN;
        rot_x(y - cu_plist[2], z - cu_plist[3], cu_plist[4], cu_plist[5], yN, zN);
        const double xx = x;
        x = yN; y = atan2(xx, zN); z = sqrt(xx * xx + zN * zN);
    }
    __device__ void param_to_xyz(const double* cu_plist, double& s, double& phi, double& r)
    {
        const double xx = r * sin(phi), yy = s, zz = r * cos(phi);
        double yN, zN;
        rot_x(yy, zz, cu_plist[4], -cu_plist[5], yN, zN);
        s = xx; phi = yN + cu_plist[2]; r = zN + cu_plist[3];
    }'''

    def __init__(self, *args, **kwargs):
        self.p = kwargs.pop('p', None)
        self.q = kwargs.pop('q', None)
        self.isCylindrical = kwargs.pop('isCylindrical', False)
        if kwargs.pop('isClosed', False):
            raise NotImplementedError('isClosed=True is outside the hot path')
        for k in ('f1', 'f2', 'parabolaAxis'):
            if kwargs.pop(k, None) is not None:
                raise NotImplementedError('{0} is resolved by xrt; give p or q'.format(k))
        if self.p is None and self.q is None:
            self.p = 10000  # the reference's default (parametric.py:434)
        if (self.p is None) == (self.q is None):
            raise ValueError('One and only one of p (or f1) or q (or f2) must be None!')
        OE.__init__(self, *args, **kwargs)
        self._reset_pq()

    def _reset_pq(self):  # parametric.py:405-430
        absPitch = _abs_pitch_to_axis(self)
        if self.p is None:
            self.y0 = self.q * np.cos(absPitch)
            self.z0 = self.q * np.sin(absPitch)
            self.parabParam = -self.q * np.sin(absPitch)**2
            gamma = absPitch
        else:
            self.y0 = -self.p * np.cos(absPitch)
            self.z0 = self.p * np.sin(absPitch)
            self.parabParam = self.p * np.sin(absPitch)**2
            gamma = -absPitch
        self.cosGamma = np.cos(gamma)
        self.sinGamma = np.sin(gamma)

    _pp2 = property(lambda self: self.parabParam**2)
    _isCyl = property(lambda self: 1. if self.isCylindrical else 0.)


ParabolicMirror = ParabolicalMirrorParam

class LauePlate(OE):
    """A flat Laue plate (the thickness is in its material): n_B = rotate_x(0, 1, -sin(alpha), -cos(alpha)) =
    (0, cos(alpha), -sin(alpha)), or (0, 1, 0) without alpha; the surface normal is (0, 0, 1).  Flat, so this is the
    built-in flat kernel with both normals given (no spliced texts)."""
    alwaysTwoNormals = True

    def _local_nB(self, second=False):
        if self.alpha:
            cosalpha, sinalpha = np.cos(self.alpha), np.sin(self.alpha)
            b, c = 0, 1
            return 0., (-sinalpha)*b - (-cosalpha)*c, (-cosalpha)*b + (-sinalpha)*c  # rotate_x(b, c, -sin, -cos)
        return 0., 1., 0.


class BentLaueCylinder(OE):
    """Cylindrically bent crystal in Laue geometry (duMond): the surface of JohannCylinder, the planes' normal turned
    by 90 degrees (+ alpha) from the surface normal.  The volumetric-diffraction model of the reference
    (local_n_depth) is outside the hot path."""
    cu_two_normals = True
    cu_plist = ('crossSectionInt', 'R', '_R2', '_cosalpha', '_sinalpha', '_hasAlpha')
    cu_local_z = '''
    __device__ double local_z(const double* cu_plist, double x, double y)
    {
        if (cu_plist[0] == 0.0) return cu_plist[1] - sqrt(cu_plist[2] - y * y);
        return y * y / 2.0 / cu_plist[1];
    }'''
    cu_local_n = _ROT + '''
    __device__ void local_n(const double* cu_plist, double x, double y, double* n)
    {
        const double R = cu_plist[1];
        double b = -y / R, c;
        if (cu_plist[0] == 0.0) {
            c = sqrt(cu_plist[2] - y * y) / R;
        } else {
            const double norm = sqrt(b * b + 1.0);
            b = b / norm;
            c = 1.0 / norm;
        }
        n[0] = 0.0; n[3] = 0.0; n[4] = b; n[5] = c;
        if (cu_plist[5] != 0.0) rot_x(b, c, -cu_plist[4], -cu_plist[3], n[1], n[2]);  // rotate_x(b, c, -sin, -cos)
        else { n[1] = c; n[2] = -b; }
    }'''

    def __init__(self, *args, **kwargs):
        self.R = kwargs.pop('R', 1.0e4)
        self.crossSection = kwargs.pop('crossSection', 'parabolic')
        if not (self.crossSection.startswith('circ') or self.crossSection.startswith('parab')):
            raise ValueError('unknown crossSection!')
        self.crossSectionInt = 0 if self.crossSection.startswith('circ') else 1
        OE.__init__(self, *args, **kwargs)

    _R2 = property(lambda self: self.R**2)
    _cosalpha = property(lambda self: float(np.cos(self.alpha)) if self.alpha else 1.)
    _sinalpha = property(lambda self: float(np.sin(self.alpha)) if self.alpha else 0.)
    _hasAlpha = property(lambda self: 1. if self.alpha else 0.)


# numpy's float floor_divide / remainder (npy_divmod, numpy/core/src/npymath/npy_math_internal.h.src)
_DIVMOD = '''
__device__ __forceinline__ void np_divmod(double a, double b, double& floordiv, double& mod)
{
    mod = fmod(a, b);
    double div = (a - mod) / b;
    if (mod != 0.0) {
        if ((b < 0.0) != (mod < 0.0)) { mod += b; div -= 1.0; }
    } else {
        mod = copysign(0.0, b);
    }
    if (div != 0.0) {
        floordiv = floor(div);
        if (div - floordiv > 0.5) floordiv += 1.0;
    } else {
        floordiv = copysign(0.0, a / b);
    }
}
'''


class BlazedGrating(OE):
    """Grating of triangular profile given by the *blaze* and *antiblaze* angles, constant line density *rho* (1/mm).
    Meant for wave propagation (the diffraction comes from the surface itself): the material is a mirror.  Like the
    reference it carries its OWN intersection method -- the first facet a ray meets, so shadowing is respected --
    instead of the bracketed root search (oes/gratings.py:492-520)."""
    cu_plist = ('rho_1', 'tanBlaze', 'tanAntiblaze', 'sinBlaze', 'cosBlaze', 'sinAntiblaze', 'cosAntiblaze', '_den')
    cu_local_z = _DIVMOD + '''
    __device__ __forceinline__ bool on_blaze_side(const double* cu_plist, double y, double& y1, double& yL)
    {
        double fd;
        np_divmod(y, cu_plist[0], fd, yL);          // y // rho_1, y % rho_1
        const double y0 = fd * cu_plist[0];
        y1 = y0 + cu_plist[0];
        const double yC = (y1 - y0) / cu_plist[7];  // 1 + tanAntiblaze / tanBlaze
        return yL > yC;
    }
    __device__ double local_z(const double* cu_plist, double x, double y)
    {
        double y1, yL;
        return on_blaze_side(cu_plist, y, y1, yL) ? (-(y1 - y)) * cu_plist[1] : (-yL) * cu_plist[2];
    }'''
    cu_local_n = '''
    __device__ void local_n(const double* cu_plist, double x, double y, double* n)
    {
        double y1, yL;
        const bool bl = on_blaze_side(cu_plist, y, y1, yL);
        n[0] = 0.0; n[1] = bl ? -cu_plist[3] : cu_plist[5]; n[2] = bl ? cu_plist[4] : cu_plist[6];
    }'''
    cu_find_intersection = '''
    __device__ void find_intersection(const double* cu_plist, double x, double y, double z, double a, double b,
                                      double c, double& t2, double& x2, double& y2, double& z2)
    {
        const double rho_1 = cu_plist[0], tB = cu_plist[1], tA = cu_plist[2];
        const double b_c = b / c;
        const double yz = y - b_c * z;
        const double n = floor(yz / rho_1);
        const double y0 = rho_1 * n, y1 = y0 + rho_1;
        double zabl = (-tA) * (yz - y0) / (1.0 + tA * b_c);
        double zbl = tB * (yz - y1) / (1.0 - tB * b_c);
        if (zabl > 0.0) zabl = zbl - 1.0;
        if (zbl > 0.0) zbl = zabl - 1.0;
        z2 = zbl;
        y2 = b_c * (z2 - z) + y;
        t2 = (y2 - y) / b;
        x2 = x + t2 * a;
    }'''

    def __init__(self, *args, **kwargs):
        self.blaze = kwargs.pop('blaze')
        self.antiblaze = kwargs.pop('antiblaze', np.pi*0.4999)
        self.rho0 = kwargs.pop('rho', 1)
        if kwargs.get('gratingDensity') is not None:
            # (the reference cannot construct this case either: reset() assigns self.rho0 whose setter calls reset()
            # again, oes/gratings.py:377-380 / :416-420 -- RecursionError; there is nothing to pin a port on)
            raise NotImplementedError('a variable line density (ticks search) is outside the hot path')
        if self.blaze == np.pi/2 or self.antiblaze == np.pi/2:
            raise NotImplementedError('vertical facets (angle == pi/2) are outside the hot path')
        OE.__init__(self, *args, **kwargs)
        self.reset()

    def reset(self):  # gratings.py:416-441
        self.rho_1 = 1. / self.rho0
        self.sinBlaze, self.cosBlaze, self.tanBlaze = np.sin(self.blaze), np.cos(self.blaze), np.tan(self.blaze)
        self.sinAntiblaze, self.cosAntiblaze, self.tanAntiblaze = \
            np.sin(self.antiblaze), np.cos(self.antiblaze), np.tan(self.antiblaze)

    _den = property(lambda self: 1 + self.tanAntiblaze/self.tanBlaze)

    def get_grating_area_fraction(self):  # gratings.py:522-534
        tanPitch = np.tan(abs(self.pitch))
        y1 = self.rho_1 * self.tanBlaze / (self.tanBlaze + tanPitch)
        z1 = -y1 * tanPitch
        y2 = self.rho_1
        z2 = 0
        d = ((y2-y1)**2 + (z2-z1)**2)**0.5
        return d * self.rho0


LIBRARY = (BentFlatMirror, ConicalMirror, JohannCylinder, JohanssonCylinder, JohannToroid, ParabolicalMirrorParam,
           BentLaueCylinder, BlazedGrating)


def prebuild(verbose=False):
    """Compile the plug-in objects of the classes above (what __graft_entry__.build() does, so that a fresh GPU box
    finds them next to libxrt_b200.so)."""
    from concurrent.futures import ThreadPoolExecutor
    from . import custom
    jobs = [dict(cu_local_z=c.cu_local_z, cu_local_n=c.cu_local_n, cu_xyz_param=c.cu_xyz_param,
                 cu_find_intersection=c.cu_find_intersection) for c in LIBRARY]
    # figure-error variants of the built-in flat and toroidal surfaces (xrt_b200/figure_error.py)
    jobs += [dict(builtin_kind=0, figure_error=True), dict(builtin_kind=1, figure_error=True),
             dict(builtin_kind=2, figure_error=True), dict(builtin_kind=3, figure_error=True),
             dict(cu_local_z=BentFlatMirror.cu_local_z, cu_local_n=BentFlatMirror.cu_local_n, figure_error=True)]
    with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
        return list(ex.map(lambda j: custom.build_plugin(verbose=verbose, **j), jobs))
