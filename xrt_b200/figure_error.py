# -*- coding: utf-8 -*-
"""Figure errors for the device elements: xrt's `OE(figureError=...)` contract
(xrt/backends/raycing/figure_error.py:45-265, hooks in oes/base.py:681-686, :744-770, :825-829, oes/reflect.py:767-776).

    GaussianBump(bumpHeight, cX, cY, sigmaX, sigmaY, limPhysX=, limPhysY=, gridStep=, baseFE=)   figure_error.py:623-715
    Waviness(amplitude, xWaveLength, yWaveLength, ...)                                           figure_error.py:717-773
    RandomRoughness(rms, rmsKind, corrLength, seed, ...)                                         figure_error.py:463-622
    FigureErrorMap(x1d, y1d, z_nm)      a given height map (what FigureErrorImported holds after reading its file)

The height map (nm) is generated once on the host exactly like the reference (its grid rules included) and turned into
the bicubic B-spline of scipy's RectBivariateSpline -- the primitive the reference itself evaluates per ray.  Knots and
coefficients go to the GPU; the per-ray evaluation of z and its two slopes (FITPACK's fpbisp / parder operation order)
is compiled into the figure-error plug-in variant of the element's reflect kernels (xrt_b200/custom.py), where the
height enters every solver iterate and the slopes turn the surface normal.  Nothing is evaluated on the CPU per ray."""
import numpy as np
import torch

maxFeHalfSize = 100  # figure_error.py:43


class FigureErrorBase(object):
    def __init__(self, name='', baseFE=None, limPhysX=None, limPhysY=None, gridStep=0.5, **kwargs):
        self.name = name
        self.baseFE = baseFE
        self.gridStep = gridStep
        self.splineOrder = 3
        self.xShift = 0.
        self.yShift = 0.
        self.limPhysX = list(limPhysX) if limPhysX is not None else [-maxFeHalfSize, maxFeHalfSize]
        self.limPhysY = list(limPhysY) if limPhysY is not None else [-maxFeHalfSize, maxFeHalfSize]
        self._dev = {}
        self.build_spline()

    # ------------------------------------------------------------------ grids (figure_error.py:167-181, :267-268)
    @staticmethod
    def next_pow2(n):
        return 1 << int(np.ceil(np.log2(n)))

    def get_grids(self):
        xlength = np.abs(self.limPhysX[-1] - self.limPhysX[0])
        ylength = np.abs(self.limPhysY[-1] - self.limPhysY[0])
        self.nx = max(self.next_pow2(xlength / self.gridStep), 128)
        self.ny = max(self.next_pow2(ylength / self.gridStep), 128)
        self.x1d = np.linspace(min(self.limPhysX), max(self.limPhysX), self.nx)
        self.y1d = np.linspace(min(self.limPhysY), max(self.limPhysY), self.ny)
        self.x2d, self.y2d = np.meshgrid(self.x1d, self.y1d)

    def generate_profile(self):
        """Height map in nm on (y1d, x1d); overridden in the subclasses."""
        self.get_grids()
        return np.zeros_like(self.x2d)

    def _base_profile(self):
        if self.baseFE is not None:
            return self.baseFE.local_z_distorted(self.x2d, self.y2d) * 1e6
        return np.zeros_like(self.x2d)

    def build_spline(self):
        from scipy import interpolate
        z = self.generate_profile()
        self.local_z_spline = interpolate.RectBivariateSpline(self.y1d, self.x1d, z, kx=self.splineOrder,
                                                              ky=self.splineOrder)
        self._dev = {}

    # host evaluation on the SMALL generation grid only (composite maps: baseFE); rays never come through here
    def local_z_distorted(self, x, y):
        x, y = np.asarray(x, dtype=float), np.asarray(y, dtype=float)
        return self.local_z_spline.ev((y + self.yShift).ravel(), (x + self.xShift).ravel()).reshape(x.shape) * 1e-6

    def get_rms(self):
        z = self.local_z_distorted(self.x2d, self.y2d) * 1e6
        return float(np.sqrt(((z - z.mean())**2).mean()))

    # ------------------------------------------------------------------ device side
    def device_spline(self, device):
        """(tu, tv, c) CUDA tensors + sizes: knots along y, knots along x, coefficients (ny_knots-4, nx_knots-4)."""
        key = str(device)
        if key not in self._dev:
            tu, tv = self.local_z_spline.get_knots()
            c = self.local_z_spline.get_coeffs()
            if self.splineOrder != 3 or len(c) != (len(tu) - 4) * (len(tv) - 4):
                raise NotImplementedError('the device evaluates bicubic figure-error splines')
            t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(device)  # noqa: E731
            self._dev[key] = (t(tu), t(tv), t(c), len(tu), len(tv))
        return self._dev[key]

    def fill_params(self, p, device):
        tu, tv, c, nu, nv = self.device_spline(device)
        p.fe_tu, p.fe_tv, p.fe_c = tu.data_ptr(), tv.data_ptr(), c.data_ptr()
        p.fe_nu, p.fe_nv = nu, nv
        p.fe_ushift, p.fe_vshift = float(self.yShift), float(self.xShift)


class FigureErrorMap(FigureErrorBase):
    """A given height map z_nm[len(y1d), len(x1d)] in nm on the local (x, y) grid."""

    def __init__(self, x1d, y1d, z_nm, **kwargs):
        self._x1d, self._y1d = np.asarray(x1d, dtype=float), np.asarray(y1d, dtype=float)
        self._z = np.asarray(z_nm, dtype=float)
        kwargs.setdefault('limPhysX', [self._x1d.min(), self._x1d.max()])
        kwargs.setdefault('limPhysY', [self._y1d.min(), self._y1d.max()])
        FigureErrorBase.__init__(self, **kwargs)

    def generate_profile(self):
        self.x1d, self.y1d = self._x1d, self._y1d
        self.x2d, self.y2d = np.meshgrid(self.x1d, self.y1d)
        return self._z + self._base_profile()


class FigureErrorSpline(FigureErrorBase):
    """A ready scipy RectBivariateSpline(y1d, x1d, z_nm, kx=3, ky=3) with its shifts -- what any of xrt's own
    FigureError objects holds as `local_z_spline`, `xShift`, `yShift` (xrt_b200.bridge converts them this way)."""

    def __init__(self, spline, xShift=0., yShift=0., name=''):
        self.name = name
        self.baseFE = None
        self.splineOrder = 3
        self.xShift, self.yShift = float(xShift), float(yShift)
        self.local_z_spline = spline
        self._dev = {}


class RandomRoughness(FigureErrorBase):
    """Stochastic roughness (figure_error.py:463-622): seeded white noise (numpy default_rng) low-passed with a Gaussian
    spectral filter of the given correlation length, normalised to an RMS height [nm] (rmsKind='height') or RMS
    slope [urad] (rmsKind='slope'; a (pitch, roll) pair gives directional slopes)."""

    def __init__(self, rms=1., rmsKind='height', corrLength=5., seed=None, **kwargs):
        self.rms, self.rmsKind, self.corrLength = rms, rmsKind, corrLength
        self.seed = np.random.SeedSequence().entropy if seed is None else seed
        kwargs['name'] = kwargs.get('name', 'random roughness')
        FigureErrorBase.__init__(self, **kwargs)

    def get_grids(self):
        FigureErrorBase.get_grids(self)
        self.dx = np.abs(self.limPhysX[-1] - self.limPhysX[0]) / self.nx  # figure_error.py:173-174
        self.dy = np.abs(self.limPhysY[-1] - self.limPhysY[0]) / self.ny

    def generate_profile(self):
        rng = np.random.default_rng(self.seed)
        self.get_grids()
        base_z = self._base_profile()
        directional = isinstance(self.rms, (tuple, list))
        z = rng.normal(loc=0.0, scale=1.0, size=(self.ny, self.nx))
        if self.corrLength is not None:  # Gaussian low-pass in the spatial-frequency domain
            spectrum = np.fft.rfft2(z)
            kx = 2 * np.pi * np.fft.rfftfreq(self.nx, d=self.dx)
            ky = 2 * np.pi * np.fft.fftfreq(self.ny, d=self.dy)
            KX, KY = np.meshgrid(kx, ky, indexing='xy')
            corrY = self.corrLength
            corrX = self.corrLength * self.rms[0] / self.rms[1] if directional else self.corrLength
            z = np.fft.irfft2(spectrum * np.exp(-0.5*(KX**2*corrX**2 + KY**2*corrY**2)), s=(self.ny, self.nx))
        z -= z.mean()
        if self.rmsKind == 'height':
            current = np.sqrt((z**2).mean())
            if current > 0:
                z *= (self.rms / current)
        elif self.rmsKind == 'slope':
            a2d, b2d = np.gradient(z*1e-6, self.y1d, self.x1d)
            rms_pitch = np.sqrt((np.arctan(a2d)**2).mean())
            rms_roll = np.sqrt((np.arctan(b2d)**2).mean())
            if directional:
                scale_y = self.rms[0] * 1e-6 / rms_pitch
                scale_x = self.rms[1] * 1e-6 / rms_roll
                spectrum = np.fft.rfft2(z)
                spectrum *= np.sqrt((scale_x * KX)**2 + (scale_y * KY)**2) / np.sqrt(KX**2 + KY**2 + 1e-30)
                z = np.fft.irfft2(spectrum, s=(self.ny, self.nx))
            else:
                z *= self.rms * 1e-6 / np.sqrt(0.5 * (rms_pitch**2 + rms_roll**2))
        else:
            raise ValueError("rmsKind is 'height' or 'slope'")
        return z + base_z


class GaussianBump(FigureErrorBase):
    def __init__(self, bumpHeight=10., cX=0., cY=0., sigmaX=10., sigmaY=10., **kwargs):
        self.bumpHeight, self.cX, self.cY, self.sigmaX, self.sigmaY = bumpHeight, cX, cY, sigmaX, sigmaY
        kwargs['name'] = kwargs.get('name', 'gaussian bump')
        FigureErrorBase.__init__(self, **kwargs)

    def generate_profile(self):  # figure_error.py:705-715
        self.get_grids()
        z = self.bumpHeight * np.exp(-(self.x2d-self.cX)**2/self.sigmaX**2 - (self.y2d-self.cY)**2/self.sigmaY**2)
        return z + self._base_profile()


class Waviness(FigureErrorBase):
    def __init__(self, amplitude=10., xWaveLength=20., yWaveLength=50., **kwargs):
        self.amplitude, self.xWaveLength, self.yWaveLength = amplitude, xWaveLength, yWaveLength
        kwargs['name'] = kwargs.get('name', 'waviness')
        FigureErrorBase.__init__(self, **kwargs)

    def generate_profile(self):  # figure_error.py:764-773
        self.get_grids()
        z = self.amplitude * np.cos(2*np.pi*self.x2d/self.xWaveLength) * np.cos(2*np.pi*self.y2d/self.yWaveLength)
        return z + self._base_profile()
