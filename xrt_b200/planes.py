# -*- coding: utf-8 -*-
"""Plane elements with xrt's signatures: `Screen` (xrt/backends/raycing/screens.py:32-302) and
`RectangularAperture` (xrt/backends/raycing/apertures.py:52-413).  Host side: orientation vectors and blade
bookkeeping; per-ray work: xrt_b200/csrc/rays.cu."""
import ctypes

import numpy as np
import torch

from . import _lib
from .beams import Beam
from .oes import REDUCTION_SLOTS, RS_FOOT

_BLADE_ORDER = ('left', 'right', 'bottom', 'top')


class xrt_plane_params(ctypes.Structure):
    """struct xrt_plane_params of include/xrt_b200.h"""
    _fields_ = [('center', ctypes.c_double * 3), ('ex', ctypes.c_double * 3), ('ey', ctypes.c_double * 3),
                ('ez', ctypes.c_double * 3), ('sinAz', ctypes.c_double), ('cosAz', ctypes.c_double),
                ('basis_mode', ctypes.c_int), ('lostNum', ctypes.c_int), ('only_positive_path', ctypes.c_int),
                ('has_compressX', ctypes.c_int), ('has_compressZ', ctypes.c_int),
                ('compressX', ctypes.c_double), ('compressZ', ctypes.c_double),
                ('n_blades', ctypes.c_int), ('blade_kind', ctypes.c_int * 4), ('blade_pos', ctypes.c_double * 4),
                ('is_beam_stop', ctypes.c_int)]


def xyz_from_xz(bl, x=None, z=None):
    """raycing.xyz_from_xz (beamline.py:290-318)."""
    if isinstance(x, (list, tuple, np.ndarray)):
        norm = sum([xc**2 for xc in x])**0.5
        retx = [xc/norm for xc in x]
    else:
        retx = (1, 0, 0.) if bl is None else (bl.cosAzimuth, -bl.sinAzimuth, 0.)
    if isinstance(z, (list, tuple, np.ndarray)):
        norm = sum([zc**2 for zc in z])**0.5
        retz = [zc/norm for zc in z]
    else:
        retz = 0., 0., 1.
    if abs(np.dot(retx, retz)) > 1e-8:
        print('x and z must be orthogonal')
    rety = np.cross(retz, retx)
    return [retx, rety, retz]


class _Plane(object):
    def _base_params(self):
        p = xrt_plane_params()
        for i in range(3):
            p.center[i] = float(self.center[i])
            p.ex[i] = float(self.x[i])
            p.ey[i] = float(self.y[i])
            p.ez[i] = float(self.z[i])
        p.basis_mode = 1
        p.sinAz = float(self.bl.sinAzimuth) if self.bl is not None else 0.
        p.cosAz = float(self.bl.cosAzimuth) if self.bl is not None else 1.
        p.lostNum = int(self.lostNum)
        return p

    def _red(self, dev):
        if not hasattr(self, '_redbuf'):
            self._redbuf = {}
        key = str(dev)
        if key not in self._redbuf:
            self._redbuf[key] = torch.zeros(REDUCTION_SLOTS, dtype=torch.int64, device=dev)
        return self._redbuf[key]

    def _decode_minmax(self, red):
        lib = _lib.load()
        vals = [lib.xrt_reduction_decode(ctypes.c_uint64(int(v) & 0xFFFFFFFFFFFFFFFF)) for v in
                red[RS_FOOT:RS_FOOT + 6].tolist()]
        return np.array(vals).reshape(2, 3)

    def local_to_global(self, x=0, y=0, z=0):
        xglo = self.center[0] + x*self.x[0] + y*self.y[0] + z*self.z[0]
        yglo = self.center[1] + x*self.x[1] + y*self.y[1] + z*self.z[1]
        zglo = self.center[2] + x*self.x[2] + y*self.y[2] + z*self.z[2]
        return xglo, yglo, zglo


class Screen(_Plane):
    """Flat screen (screens.py:32-302)."""

    def __init__(self, bl=None, name='', center=[0, 0, 0], x='auto', z='auto', compressX=None, compressZ=None,
                 **kwargs):
        self.bl = bl
        self.name = name
        if bl is not None:
            if self not in bl.screens:
                bl.screens.append(self)
            self.ordinalNum = len(bl.screens)
        else:
            self.ordinalNum = 1
        self.lostNum = -self.ordinalNum - 2000  # screens.py:76
        self.center = list(center)
        self.x, self.y, self.z = xyz_from_xz(bl, None if isinstance(x, str) else x,
                                             None if isinstance(z, str) else z)
        self.compressX = compressX
        self.compressZ = compressZ
        self.footprint = []

    def _expose_params(self, onlyPositivePath=False):
        p = self._base_params()
        p.only_positive_path = int(bool(onlyPositivePath))
        if self.compressX:
            p.has_compressX, p.compressX = 1, float(self.compressX)
        if self.compressZ:
            p.has_compressZ, p.compressZ = 1, float(self.compressZ)
        return p

    def _fused_target(self, beam, onlyPositivePath=False):
        """(plane parameters, output Beam, reduction slots) for an optical element whose kernel continues with this
        screen's expose (OE.reflect(exposeOn=screen))."""
        return self._expose_params(onlyPositivePath), beam.like(), self._red(beam.device)

    def expose(self, beam=None, onlyPositivePath=False, withHistogram=False, withFootprint=False):
        """Returns the beam in the local system of the screen (screens.py:225-302)."""
        if withHistogram:
            raise NotImplementedError('histogramming is the next row of the scope table (SURVEY.md 8f1)')
        lib = _lib.load()
        p = self._expose_params(onlyPositivePath)
        blo = beam.like()
        red = self._red(beam.device)
        cin, cout = beam.c_struct(), blo.c_struct()
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(lib.xrt_expose_f64(len(beam), ctypes.byref(p), ctypes.byref(cin), ctypes.byref(cout),
                                      red.data_ptr(), stream), 'xrt_expose_f64')
        if withFootprint:
            self.footprint = [self._decode_minmax(red)]
        blo.parentId = getattr(self, 'uuid', self.name)
        return blo


class HemisphericScreen(Screen):
    """Hemispheric screen (screens.py:422-545): rays are exposed on the sphere of radius *R* around *center*; the
    returned local beam carries *theta* and *phi* (minus their offsets).  Default orientation: local x along the beam
    line, local z horizontal (screens.py:470-487)."""

    def __init__(self, bl=None, name='', center=[0, 0, 0], R=1000., x='auto', z='auto', phiOffset=0, thetaOffset=0,
                 **kwargs):
        Screen.__init__(self, bl, name, center, **kwargs)
        self.R = R
        self.phiOffset = phiOffset
        self.thetaOffset = thetaOffset
        sinAz = bl.sinAzimuth if bl is not None else 0.
        cosAz = bl.cosAzimuth if bl is not None else 1.
        if isinstance(x, (list, tuple, np.ndarray)):
            norm = sum([xc**2 for xc in x])**0.5
            self.x = [xc/norm for xc in x]
        else:
            self.x = [sinAz, cosAz, 0.]
        if isinstance(z, (list, tuple, np.ndarray)):
            norm = sum([zc**2 for zc in z])**0.5
            self.z = [zc/norm for zc in z]
        else:
            self.z = [cosAz, -sinAz, 0.]
        self.y = list(np.cross(self.z, self.x))

    def expose(self, beam=None, onlyPositivePath=False):
        lib = _lib.load()
        p = self._base_params()
        p.only_positive_path = int(bool(onlyPositivePath))
        blo = beam.like()
        blo.theta = torch.empty(len(beam), dtype=torch.float64, device=beam.device)
        blo.phi = torch.empty(len(beam), dtype=torch.float64, device=beam.device)
        cin, cout = beam.c_struct(), blo.c_struct()
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(lib.xrt_expose_hemispheric_f64(len(beam), ctypes.byref(p), float(self.R), float(self.thetaOffset),
                                                  float(self.phiOffset), ctypes.byref(cin), ctypes.byref(cout),
                                                  stream), 'xrt_expose_hemispheric_f64')
        blo.parentId = getattr(self, 'uuid', self.name)
        return blo

    def prepare_wave(self, *a, **k):
        raise NotImplementedError('HemisphericScreen.prepare_wave is outside the hot path')


def _screen_prepare_wave(self, prevOE, dim1, dim2, dy=0, rw=None, condition=None):
    """Screen.prepare_wave (screens.py:304-365): a mesh of receiving points on the screen for waves.diffract or
    source.shine(wave=...)."""
    from . import waves
    return waves.screen_prepare_wave(self, prevOE, dim1, dim2, dy=dy, condition=condition)


Screen.prepare_wave = _screen_prepare_wave


class RectangularAperture(_Plane):
    """Rectangular aperture / slit (apertures.py:52-413)."""

    #: the reference's default blades (apertures.py:64-65): a script that gives neither kind / opening nor blades gets
    #: a +-10 mm opening, not an open plane
    DEFAULT_BLADES = {'left': -10, 'right': 10, 'bottom': -10, 'top': 10}

    def __init__(self, bl=None, name='', center=[0, 0, 0], kind=None, opening=None, x='auto', z='auto',
                 alarmLevel=None, blades=DEFAULT_BLADES, **kwargs):
        self.bl = bl
        self.name = name
        if bl is not None:
            if self not in bl.slits:
                bl.slits.append(self)
            self.ordinalNum = len(bl.slits)
        else:
            self.ordinalNum = 1
        self.lostNum = -self.ordinalNum - 1000  # apertures.py:131
        self.center = list(center)
        self.x, self.y, self.z = xyz_from_xz(bl, None if isinstance(x, str) else x,
                                             None if isinstance(z, str) else z)
        self.xyz = [self.x, self.y, self.z]
        self.limOptX = list(kwargs.get('limOptX', [-500, 500]))
        self.limOptY = list(kwargs.get('limOptY', [-500, 500]))
        self.isBeamStop = kwargs.get('isBeamStop', False)
        self.spotLimits = [0, 0, 0, 0]
        self.alarmLevel = alarmLevel
        if kind is not None or opening is not None:  # kind / opening take precedence over blades (apertures.py:162-174)
            if kind is None and opening is not None:
                kind = list(_BLADE_ORDER[:len(opening)]) if hasattr(opening, '__len__') else [_BLADE_ORDER[0]]
            if isinstance(kind, str):
                kind, opening = (kind,), [opening]
            self.blades = {k: v for k, v in zip(kind or [], opening or []) if v is not None}
        elif blades is not None:
            self.blades = {k: v for k, v in dict(blades).items() if v is not None}
        else:
            self.blades = {}
        self.set_optical_limits()

    def set_optical_limits(self):
        for akind, d in self.blades.items():  # apertures.py:289-301
            td = float(d)
            if akind.startswith('l'):
                self.limOptX[0] = td
            elif akind.startswith('r'):
                self.limOptX[1] = td
            elif akind.startswith('b'):
                self.limOptY[0] = td
            elif akind.startswith('t'):
                self.limOptY[1] = td

    def propagate(self, beam=None, needNewGlobal=False, withSpotLimits=False):
        """Returns the local beam; rays intercepted by a blade get state = lostNum in the returned beam and
        in *beam* itself, like the reference (apertures.py:333-413)."""
        if needNewGlobal:
            raise NotImplementedError('needNewGlobal=True is not on the hot path')
        lib = _lib.load()
        p = self._base_params()
        kinds = {'l': 0, 'r': 1, 'b': 2, 't': 3}
        items = self._blade_items() if hasattr(self, '_blade_items') else \
            [(kinds[k[0]], float(d)) for k, d in self.blades.items() if k[0] in kinds]
        p.n_blades = len(items)
        for i, (k, d) in enumerate(items):
            p.blade_kind[i] = k
            p.blade_pos[i] = d
        p.is_beam_stop = int(bool(self.isBeamStop))
        lo = beam.like()
        red = self._red(beam.device)
        cin, cout = beam.c_struct(), lo.c_struct()
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(lib.xrt_propagate_rect_f64(len(beam), ctypes.byref(p), ctypes.byref(cin), ctypes.byref(cout),
                                              red.data_ptr(), stream), 'xrt_propagate_rect_f64')
        if withSpotLimits:
            mm = self._decode_minmax(red)
            if np.isfinite(mm).all():
                self.spotLimits = [min(self.spotLimits[0], mm[0, 0]), max(self.spotLimits[1], mm[1, 0]),
                                   min(self.spotLimits[2], mm[0, 2]), max(self.spotLimits[3], mm[1, 2])]
        lo.parentId = getattr(self, 'uuid', self.name)
        return lo


class RoundAperture(RectangularAperture):
    """Round aperture: a pipe or a flange (apertures.py:668-875).  The same kernel as the rectangular one with a single
    'blade' of the round kind: rays with sqrt(x^2 + z^2) > r are lost (inverted for a beam stop)."""

    def __init__(self, bl=None, name='', center=[0, 0, 0], r=1, x='auto', z='auto', alarmLevel=None, **kwargs):
        RectangularAperture.__init__(self, bl, name, center, x=x, z=z, alarmLevel=alarmLevel, blades={},
                                     isBeamStop=kwargs.get('isBeamStop', False))
        self.r = r
        self.limOptX = [-r, r]
        self.limOptY = [-r, r]
        self.shape = 'round'

    def _blade_items(self):
        return [(4, float(self.r))]

    def prepare_wave(self, prevOE, nrays, rw=None, xy=None):
        """RoundAperture.prepare_wave (apertures.py:848-875): random receiving points in the disk (uniform in area)."""
        from . import waves
        return waves.aperture_prepare_wave(self, prevOE, nrays, xy=xy, round_r=self.r)


def _aperture_prepare_wave(self, prevOE, nrays, rw=None, xy=None):
    """RectangularAperture.prepare_wave (apertures.py:467-499): random receiving points in the opening."""
    from . import waves
    return waves.aperture_prepare_wave(self, prevOE, nrays, xy=xy)


RectangularAperture.prepare_wave = _aperture_prepare_wave
