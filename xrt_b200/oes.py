# -*- coding: utf-8 -*-
"""Optical elements of the hot path with xrt's public signatures:

    OE(bl, name, center, pitch, roll, yaw, positionRoll, ..., material, alpha, limPhysX, limOptX, limPhysY,
       limOptY, ...)                                   xrt/backends/raycing/oes/base.py:92-104
    OE.reflect(beam, needLocal=True, ...) -> (beamGlobal, beamLocal)         oes/reflect.py:18-19
    ToroidMirror(..., R, r)                                                   oes/__init__.py:320-411
    EllipticalMirrorParam(..., p, q, isCylindrical)                           oes/parametric.py:9-249
    DCM(..., bragg, cryst1roll, cryst2roll, cryst2pitch, cryst2finePitch, cryst2perpTransl,
        cryst2longTransl, limPhysX2, limPhysY2, limOptX2, limOptY2, material2)          oes/dcm.py:12-354
    DCM.double_reflect(beam, needLocal=True) -> (beamGlobal, beamLocal1, beamLocal2)    oes/dcm.py:248-250

The host side only marshals geometry into `xrt_oe_params` (include/xrt_b200.h); all per-ray work is one fused
CUDA kernel per surface (xrt_b200/csrc/reflect.cu) on device-resident Beams.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .beams import Beam
from .materials import xrt_material_params

XRT_MAX_ROT = 8
maxHalfSizeOfOE = 1000.  # raycing/__init__.py:93-94
maxDepthOfOE = 100.
REDUCTION_SLOTS = 64
RS_FOOT, RS_NGOOD, RS_NGOODN, RS_OOR = 28, 34, 35, 36


class xrt_oe_params(ctypes.Structure):
    """struct xrt_oe_params of include/xrt_b200.h"""
    _fields_ = [('center', ctypes.c_double * 3), ('sinAz', ctypes.c_double), ('cosAz', ctypes.c_double),
                ('from_global', ctypes.c_int), ('to_global', ctypes.c_int),
                ('n_rot_in', ctypes.c_int), ('rot_in_axis', ctypes.c_int * XRT_MAX_ROT),
                ('rot_in_cos', ctypes.c_double * XRT_MAX_ROT), ('rot_in_sin', ctypes.c_double * XRT_MAX_ROT),
                ('n_rot_out', ctypes.c_int), ('rot_out_axis', ctypes.c_int * XRT_MAX_ROT),
                ('rot_out_cos', ctypes.c_double * XRT_MAX_ROT), ('rot_out_sin', ctypes.c_double * XRT_MAX_ROT),
                ('shift', ctypes.c_double * 3), ('roll', ctypes.c_double),
                ('surface', ctypes.c_int), ('sp', ctypes.c_double * 8),
                ('asymmetric', ctypes.c_int), ('nB', ctypes.c_double * 3),
                ('physX', ctypes.c_double * 2), ('physY', ctypes.c_double * 2),
                ('optX', ctypes.c_double * 2), ('optY', ctypes.c_double * 2),
                ('has_optX', ctypes.c_int), ('has_optY', ctypes.c_int),
                ('over_mask', ctypes.c_int), ('lostNum', ctypes.c_int),
                ('brk_lim', (ctypes.c_double * 2) * 3),
                ('created_by_diffract', ctypes.c_int), ('zero_dead_local', ctypes.c_int),
                ('good_mode', ctypes.c_int), ('want_footprint', ctypes.c_int),
                ('no_intersection', ctypes.c_int), ('from_vacuum', ctypes.c_int), ('dead_state_lost', ctypes.c_int),
                ('grating', ctypes.c_int), ('g_ncoef', ctypes.c_int), ('g_rho0', ctypes.c_double),
                ('g_coef', ctypes.c_double * 8), ('g_order', ctypes.c_double), ('g_amp', ctypes.c_double),
                ('shape_round', ctypes.c_int), ('shape_r', ctypes.c_double * 4),
                ('fzp_rn', ctypes.c_void_p), ('fzp_n', ctypes.c_int), ('fzp_black', ctypes.c_int),
                ('multi', ctypes.c_int),
                ('fe_tu', ctypes.c_void_p), ('fe_tv', ctypes.c_void_p), ('fe_c', ctypes.c_void_p),
                ('fe_nu', ctypes.c_int), ('fe_nv', ctypes.c_int),
                ('fe_ushift', ctypes.c_double), ('fe_vshift', ctypes.c_double),
                ('multi_tmin', ctypes.c_void_p)]


_AXIS = {'x': 0, 'y': 1, 'z': 2}


def _rotation_list(rotationSequence, pitch, roll, yaw):
    """Elementary rotations of raycing.rotate_beam (_rotate.py:23-57) in application order, zero angles
    skipped, cos/sin taken with numpy exactly as the reference does."""
    angles = {'z': yaw, 'y': roll, 'x': pitch}
    if rotationSequence[0] == '-':
        seq = rotationSequence[6] + rotationSequence[4] + rotationSequence[2]
    else:
        seq = rotationSequence[1] + rotationSequence[3] + rotationSequence[5]
    out = []
    for s in seq:
        angle = angles[s]
        if angle != 0:
            out.append((_AXIS[s], float(np.cos(angle)), float(np.sin(angle))))
    return out


def _fill_rot(params, which, rots):
    if len(rots) > XRT_MAX_ROT:
        raise ValueError('too many elementary rotations')
    setattr(params, 'n_rot_' + which, len(rots))
    for i, (ax, c, s) in enumerate(rots):
        getattr(params, 'rot_{0}_axis'.format(which))[i] = ax
        getattr(params, 'rot_{0}_cos'.format(which))[i] = c
        getattr(params, 'rot_{0}_sin'.format(which))[i] = s


class OE(object):
    """Flat optical element (mirror or crystal).

    A subclass with another surface either names a built-in kernel (`surfaceKind` 1..3) or carries the CUDA texts of
    its surface -- `cu_plist`, `cu_local_z`, `cu_local_n` (+ `cu_xyz_param` for a parametric element), the
    counterparts of the reference's `cl_plist`, `cl_local_z`, `cl_local_n`, `cl_xyz_param` (oes/base.py:548-569) --
    which xrt_b200.custom compiles into the same fused kernels (surface kinds 4 / 5)."""
    surfaceKind = 0
    isParametric = False
    cu_plist = ()
    cu_local_z = None
    cu_local_n = None
    cu_xyz_param = None
    cu_find_intersection = None  # an element's own intersection method instead of the bracketed root search
    cu_two_normals = None  # None: two normals (n_B, n_s) are returned iff alpha is set, like OE.local_n

    def __init__(self, bl=None, name='', center=[0, 0, 0], pitch=0, roll=0, yaw=0, positionRoll=0,
                 rotationSequence='RzRyRx', extraPitch=0, extraRoll=0, extraYaw=0,
                 extraRotationSequence='RzRyRx', alarmLevel=None, surface=None, material=None, alpha=None,
                 limPhysX=None, limOptX=None, limPhysY=None, limOptY=None, isParametric=False, shape='rect',
                 targetOpenCL='CUDA', precisionOpenCL='float64', **kwargs):
        self.bl = bl
        self.name = name
        if bl is not None:
            if self not in bl.oes:  # ordinal numbering -> lostNum (base.py:268-272)
                bl.oes.append(self)
            self.ordinalNum = len(bl.oes)
            self.lostNum = -self.ordinalNum
        else:
            self.ordinalNum = 1
            self.lostNum = -1
        self.center = list(center)
        self.pitch = pitch
        self.roll = roll
        self.yaw = yaw
        self.positionRoll = positionRoll
        self.rotationSequence = rotationSequence
        self.extraPitch, self.extraRoll, self.extraYaw = extraPitch, extraRoll, extraYaw
        self.extraRotationSequence = extraRotationSequence
        self.alarmLevel = alarmLevel
        self.material = material
        self.alpha = alpha
        self.figureError = kwargs.pop('figureError', None)  # xrt_b200.figure_error object (base.py:160, :310)
        self.shape = shape
        if not (isinstance(shape, str) and (shape.startswith('re') or shape.startswith('ro'))):
            raise NotImplementedError("shapes 'rect' and 'round' are on the hot path (polygons are not)")
        self.overEdge = kwargs.pop('overEdge', 'yMax')
        # gratings (base.py:213-230, :491-508): gratingDensity = ['x'|'y', rho0 (1/mm), c0, c1, ...], order: int
        self.gratingDensity = kwargs.pop('gratingDensity', None)
        order = kwargs.pop('order', None)
        self.order = 1 if order is None else order
        if self.gratingDensity is not None and self.material is None:
            from .materials import EmptyMaterial
            self.material = EmptyMaterial()
        default = [-maxHalfSizeOfOE, maxHalfSizeOfOE]
        self.limPhysX = list(limPhysX) if limPhysX is not None else list(default)  # base.py:465-476
        self.limPhysY = list(limPhysY) if limPhysY is not None else list(default)
        self.limOptX = limOptX
        self.limOptY = limOptY
        self.dx = 0
        self.footprint = []
        self.targetOpenCL = targetOpenCL
        self.precisionOpenCL = precisionOpenCL
        if targetOpenCL is None:
            raise _lib.XrtB200Error('xrt_b200 elements always run on the GPU (targetOpenCL=None has no numpy '
                                    'path to fall back to)')
        self._red = {}

    # ------------------------------------------------------------------ geometry -> C struct
    def _surface_params(self):
        out = []
        for name in self.cu_plist:  # cl_plist marshalling (base.py:897-899): up to 8 numbers, sequences flattened
            v = getattr(self, name)
            out.extend([float(u) for u in v] if isinstance(v, (list, tuple, np.ndarray)) else [float(v)])
        if len(out) > 8:
            raise ValueError('cu_plist: at most 8 numbers')
        return out

    def _plugin(self):
        """The plug-in object that serves this element -- a spliced surface and / or the figure-error variant of its
        kernels -- or None for the built-in surfaces without figure error."""
        fe = self.figureError is not None
        if self.cu_local_z is None:
            if not fe:
                return None
            if self.surfaceKind not in (0, 1, 2, 3):
                raise NotImplementedError('figure errors are on the hot path for flat, toroidal, elliptic, parabolic '
                                          '(lenses) and spliced surfaces')
            if self.alpha:
                # the reference has no behaviour to match: np.asarray(oeNormal) fails on the scalar normal of the
                # planes beside the turned per-ray surface normal (reflect.py:767-790)
                raise NotImplementedError('a figure error on an asymmetrically cut crystal fails in the reference '
                                          '(reflect.py:790); outside the hot path')
            from . import custom
            return custom.plugin_for(builtin_kind=self.surfaceKind, figure_error=True)
        if self.cu_local_n is None:
            raise _lib.XrtB200Error('a spliced surface needs both cu_local_z and cu_local_n')
        from . import custom
        return custom.plugin_for(self.cu_local_z, self.cu_local_n, self.cu_xyz_param, figure_error=fe,
                                 cu_find_intersection=self.cu_find_intersection)

    def _limits(self, second=False):
        sfx = '2' if second else ''
        return (getattr(self, 'limPhysX' + sfx), getattr(self, 'limPhysY' + sfx),
                getattr(self, 'limOptX' + sfx), getattr(self, 'limOptY' + sfx))

    def _local_nB(self, second=False):
        # OE.local_n with alpha (base.py:736-739): rotate_x(b=0, c=1, cosalpha, -sinalpha); DCM.local_n2 flips y
        cosalpha, sinalpha = np.cos(self.alpha), np.sin(self.alpha)
        b, c = 0., 1.
        bAlpha, cAlpha = cosalpha*b - (-sinalpha)*c, (-sinalpha)*b + cosalpha*c
        if second:
            bAlpha *= -1
        return 0., bAlpha, cAlpha

    def _make_params(self, pitch, roll, yaw, dx=0, dy=0, dz=0, is2ndXtal=False, from_global=True,
                     to_global=True, createdByDiffract=False):
        p = xrt_oe_params()
        for i in range(3):
            p.center[i] = float(self.center[i])
        p.sinAz = float(self.bl.sinAzimuth) if self.bl is not None else 0.
        p.cosAz = float(self.bl.cosAzimuth) if self.bl is not None else 1.
        p.from_global = int(from_global)
        p.to_global = int(to_global)
        extraSign = -1. if is2ndXtal else 1.
        hasExtra = bool(self.extraPitch or self.extraRoll or self.extraYaw)
        rin = []
        if is2ndXtal:  # reflect.py:619-621
            rin += _rotation_list('RzRyRx', 0, -np.pi, 0)
        rin += _rotation_list(self.rotationSequence, -pitch, -roll, -yaw)
        if hasExtra:
            rin += _rotation_list(self.extraRotationSequence, -extraSign*self.extraPitch, -self.extraRoll,
                                  -extraSign*self.extraYaw)
        rout = []
        if hasExtra:
            rout += _rotation_list('-'+self.extraRotationSequence, extraSign*self.extraPitch, self.extraRoll,
                                   extraSign*self.extraYaw)
        rout += _rotation_list('-'+self.rotationSequence, pitch, roll, yaw)
        if is2ndXtal:
            rout += _rotation_list('RzRyRx', 0, np.pi, 0)
        _fill_rot(p, 'in', rin)
        _fill_rot(p, 'out', rout)
        p.shift[0], p.shift[1], p.shift[2] = float(dx or 0), float(dy or 0), float(dz or 0)
        p.roll = float(roll)
        p.surface = self.surfaceKind
        if self.cu_local_z is not None:
            p.surface = 5 if self.cu_xyz_param else 4
        for i, v in enumerate(self._surface_params()):
            p.sp[i] = float(v)
        p.asymmetric = 1 if (self.alpha or getattr(self, 'alwaysTwoNormals', False)) else 0
        if self.cu_local_z is not None and self.cu_two_normals is not None:
            p.asymmetric = int(bool(self.cu_two_normals))
        if p.asymmetric and self.cu_local_z is None:
            nB = self._local_nB(is2ndXtal)
            for i in range(3):
                p.nB[i] = nB[i]
        physX, physY, optX, optY = self._limits(is2ndXtal)
        p.physX[0], p.physX[1] = float(physX[0]), float(physX[1])
        p.physY[0], p.physY[1] = float(physY[0]), float(physY[1])
        p.has_optX = int(optX is not None)
        p.has_optY = int(optY is not None)
        if optX is not None:
            p.optX[0], p.optX[1] = float(optX[0]), float(optX[1])
        if optY is not None:
            p.optY[0], p.optY[1] = float(optY[0]), float(optY[1])
        ovE = str(self.overEdge).lower()
        p.over_mask = (1 if 'xmin' in ovE else 0) | (2 if 'xmax' in ovE else 0) | \
            (4 if 'ymin' in ovE else 0) | (8 if 'ymax' in ovE else 0)
        p.lostNum = int(self.lostNum)
        if self.shape.startswith('ro'):  # base.py:1140-1156
            if is2ndXtal:
                raise NotImplementedError("shape='round' on a second crystal")
            centerX = (physX[0]+physX[1]) * 0.5
            centerX = 0 if np.isnan(centerX) else centerX
            radiusX = (physX[1]-physX[0]) * 0.5
            centerY = (physY[0]+physY[1]) * 0.5
            centerY = 0 if np.isnan(centerY) else centerY
            radiusY = (physY[1]-physY[0]) * 0.5
            if not np.isinf(radiusX):
                p.shape_round = 1
                for i, v in enumerate((centerX, radiusX, centerY, radiusY)):
                    p.shape_r[i] = float(v)
            else:
                p.shape_round = 1  # infinite ellipse: nothing is lost
                for i, v in enumerate((0., 1e300, 0., 1e300)):
                    p.shape_r[i] = v
        # _set_t limits (base.py:1233-1239)
        for ax, lim, dflt in ((0, physX, maxHalfSizeOfOE), (1, physY, maxHalfSizeOfOE), (2, None, maxDepthOfOE)):
            if lim is None:
                lo, hi = -dflt, dflt
            else:
                lo = lim[0] if lim[0] > -np.inf else -dflt
                hi = lim[1] if lim[1] < np.inf else dflt
            p.brk_lim[ax][0], p.brk_lim[ax][1] = float(lo), float(hi)
        p.created_by_diffract = int(createdByDiffract)
        p.want_footprint = 1
        p.from_vacuum = 1
        self._fill_grating(p, self.material2 if (is2ndXtal and hasattr(self, 'material2')) else self.material)
        if self.figureError is not None:
            self.figureError.fill_params(p, torch.device('cuda', torch.cuda.current_device()))
        return p

    def _fill_grating(self, p, material):
        """toWhere = 3 (reflect.py:743-744): the material's kind decides; 'auto' becomes 'grating' when the OE has a
        gratingDensity (base.py:1089-1093).  OE.local_g (base.py:688-717) is marshalled as polynomial coefficients."""
        p.g_amp = -1.
        kind = getattr(material, 'kind', None)
        if kind == 'auto' and self.gratingDensity is not None:
            kind = 'grating'
        if hasattr(self, 'rn') and kind in ('FZP', 'auto'):  # NormalFZP: assign_auto_material_kind -> 'FZP'
            if not isinstance(self.order, (int, np.integer)):
                raise NotImplementedError('give OE.order as one int')
            dev = torch.device('cuda', torch.cuda.current_device())
            key = str(dev)
            if getattr(self, '_rn_dev', (None,))[0] != key:
                self._rn_dev = (key, torch.from_numpy(np.ascontiguousarray(self.rn, dtype=np.float64)).to(dev))
            p.grating = 4
            p.g_order = float(self.order)
            p.fzp_rn = self._rn_dev[1].data_ptr()
            p.fzp_n = len(self.rn)
            p.fzp_black = int(bool(self.isCentralZoneBlack))
            return
        if kind != 'grating':
            return
        if not isinstance(self.order, (int, np.integer)):
            raise NotImplementedError('a sequence of diffraction orders (random choice per ray) is outside the '
                                      'hot path; give OE.order as one int')
        p.g_order = float(self.order)
        rhoList = self.gratingDensity
        if rhoList is not None:
            coefs = list(rhoList[2:])
            if len(coefs) > 8:
                raise ValueError('gratingDensity: at most 8 polynomial coefficients')
            p.grating = 1 if rhoList[0] == 'x' else 2
            p.g_rho0 = float(rhoList[1])
            p.g_ncoef = len(coefs)
            for i, c in enumerate(coefs):
                p.g_coef[i] = float(c)
        else:
            p.grating = 3
            p.g_rho0 = -100.  # default local_g: (0, rho=-100., 0) (base.py:688, :717)
        eff = getattr(material, 'efficiency', None)
        if eff is not None:  # Material.get_grating_efficiency (material.py:391-413), constant efficiencies
            resI = 0.
            for o, v in eff:
                if o == self.order:
                    resI = v
            p.g_amp = float(resI)**0.5

    # ------------------------------------------------------------------ kernel launch
    def _launch(self, params, material, beam, gb, lb, fallback=None, index_offset=0, allreduce=None, phases=None,
                screen=None):
        """One _reflect_local pass.  *screen* = (plane parameters, screen Beam, the screen's reduction slots): the main
        pass continues with Screen.expose in the same kernel and writes the screen beam instead of *gb* (None)."""
        lib = _lib.load()
        dev = beam.device
        key = str(dev)
        if key not in self._red:
            self._red[key] = torch.zeros(REDUCTION_SLOTS, dtype=torch.int64, device=dev)
        red = self._red[key]
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        if params.grating and (params.g_amp >= 0 or params.grating == 4 or not hasattr(material, 'device_tables')):
            # ... or a zone plate: kind 'FZP' has amplitude (1, 1) (material.py:457-458)
            # EmptyMaterial or a tabulated efficiency: no Fresnel amplitude (reflect.py:1005-1011): kind 7
            mat = xrt_material_params()
            mat.kind = 7
            mat_ref, tab_ptr, tlen = ctypes.byref(mat), None, 0
        elif material is not None:
            tab, tlen, offs, lens = material.device_tables(dev)
            mat = material.c_struct(offs, lens)
            if params.grating:
                mat.kind = 6  # coated grating: grating deflection + mirror amplitude
            mat_ref, tab_ptr = ctypes.byref(mat), tab.data_ptr()
        else:
            mat_ref, tab_ptr, tlen = None, None, 0
        cin = beam.c_struct()
        cgb = gb.c_struct() if gb is not None else None
        clb = lb.c_struct() if lb is not None else None
        cfb = fallback.c_struct() if fallback is not None else None
        args = (len(beam), ctypes.byref(params), mat_ref, tab_ptr, tlen, ctypes.byref(cin),
                ctypes.byref(cfb) if cfb is not None else None, ctypes.byref(cgb) if cgb is not None else None,
                ctypes.byref(clb) if clb is not None else None, red.data_ptr())
        profile = getattr(self, 'profileEvents', None)
        plugin = self._plugin()
        if screen is not None:
            if plugin is not None:
                raise NotImplementedError('the fused reflect + expose kernel serves the built-in surfaces')
            sp, sbeam, sred = screen
            csb = sbeam.c_struct()
            fused = lib.xrt_reflect_expose_f64

            def xrt_reflect_f64(n, oe, mat, tab, tl, cin_, cfb_, cgb_, clb_, red_, ph, off, st):
                return fused(n, oe, mat, tab, tl, cin_, cfb_, clb_, red_, ctypes.byref(sp), ctypes.byref(csb),
                             sred.data_ptr(), ph, off, st)
        else:
            xrt_reflect_f64 = lib.xrt_reflect_f64 if plugin is None else plugin.reflect
        if phases is not None:  # multiple_reflect interleaves the derivative stage between the phases
            for ph in phases:
                if allreduce is None:
                    _lib.check(xrt_reflect_f64(*args, ph, index_offset, stream), 'xrt_reflect_f64')
                    continue
                # ray shards: the beam-wide scalars are all-reduced after each pre-pass, also when the caller
                # interleaves its own stages (pre-pass A = bit 0, B = bit 2, main = bit 1)
                if ph & 1:
                    _lib.check(xrt_reflect_f64(*args, 1, index_offset, stream), 'xrt_reflect_f64 (pre-pass A)')
                    allreduce(red)
                if ph & 4:
                    _lib.check(xrt_reflect_f64(*args, 4, index_offset, stream), 'xrt_reflect_f64 (pre-pass B)')
                    allreduce(red, False)
                if ph & 2:
                    _lib.check(xrt_reflect_f64(*args, 2, index_offset, stream), 'xrt_reflect_f64 (main pass)')
        elif allreduce is None and profile is None:
            _lib.check(xrt_reflect_f64(*args, 7, index_offset, stream), 'xrt_reflect_f64')
        else:
            # the three phases as separate calls: ray shards all-reduce the beam-wide scalars (min/max/sum over
            # the 64 slots, xrt_b200/parallel.py) after each pre-pass; bench.py takes per-kernel CUDA events
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if profile is not None else None
            if evs:
                evs[0].record()
            _lib.check(xrt_reflect_f64(*args, 1, index_offset, stream), 'xrt_reflect_f64 (pre-pass A)')
            if evs:
                evs[1].record()
            if allreduce is not None:
                allreduce(red)
            if evs:
                evs.append(torch.cuda.Event(enable_timing=True))
                evs[4].record()
            _lib.check(xrt_reflect_f64(*args, 4, index_offset, stream), 'xrt_reflect_f64 (pre-pass B)')
            if evs:
                evs[2].record()
            if allreduce is not None:
                allreduce(red, False)  # counters were summed above
            if evs:
                evs.append(torch.cuda.Event(enable_timing=True))
                evs[5].record()
            _lib.check(xrt_reflect_f64(*args, 2, index_offset, stream), 'xrt_reflect_f64 (main pass)')
            if evs:
                evs[3].record()
                # kernel intervals: A = 0->1, B = 4->2, main = 5->3 (the all-reduces sit in between)
                profile.append(evs)
        return red

    def _check_tables(self, red):
        """Element.get_f1f2 raises ValueError when an energy is outside the table (element.py:254-260)."""
        vals = red[RS_OOR:RS_OOR + 3].tolist()  # out-of-table count, rays on either side of an asymmetric crystal
        if int(vals[0]) > 0:
            raise ValueError('E is out of the f1/f2 data table range')
        if int(vals[1]) > 0 and int(vals[2]) > 0:
            # reflect.py:575-576 takes ONE sign (of the mean n_B.k over the beam) for the deflection by the asymmetric
            # cut; the kernel takes it per ray, which is the same unless the beam meets the planes from both sides
            raise NotImplementedError(
                'rays meet the asymmetric crystal from both sides of its Bragg planes ({0} / {1} rays): the reference '
                'deflects all of them with the sign of the beam mean; this case is outside the hot path'.format(
                    int(vals[1]), int(vals[2])))

    def _footprint_from(self, red):
        lib = _lib.load()
        vals = [lib.xrt_reduction_decode(ctypes.c_uint64(int(v) & 0xFFFFFFFFFFFFFFFF)) for v in
                red[RS_FOOT:RS_FOOT + 6].tolist()]
        return np.array(vals).reshape(2, 3)

    def local_n_device(self, x, y, z):
        """Surface normal (nx, ny, nz) at local Cartesian points given as CUDA tensors: OE.local_n after
        xyz_to_param for parametric surfaces (waves.py:675-686)."""
        if self.figureError is not None:
            raise NotImplementedError('wave-mode helpers with a figure error are outside the hot path')
        lib = _lib.load()
        n = x.numel()
        out = torch.empty((3, n), dtype=torch.float64, device=x.device)
        params = self._make_params(self.pitch, self.roll + self.positionRoll, self.yaw, self.dx)
        plugin = self._plugin()
        oe_normals = lib.xrt_oe_normals_f64 if plugin is None else plugin.oe_normals
        rc = oe_normals(n, ctypes.byref(params), x.contiguous().data_ptr(), y.contiguous().data_ptr(),
                        z.contiguous().data_ptr(), out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(),
                        ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, 'xrt_oe_normals_f64')
        return out[0], out[1], out[2]

    def local_z_device(self, x, y):
        """Surface height at local (x, y) CUDA tensors, as OE.prepare_wave takes it (oes/reflect.py:335-340)."""
        lib = _lib.load()
        n = x.numel()
        out = torch.empty(n, dtype=torch.float64, device=x.device)
        params = self._make_params(self.pitch, self.roll + self.positionRoll, self.yaw, self.dx)
        plugin = self._plugin()
        oe_surface = lib.xrt_oe_surface_f64 if plugin is None else plugin.oe_surface
        rc = oe_surface(n, ctypes.byref(params), x.contiguous().data_ptr(), y.contiguous().data_ptr(),
                        out.data_ptr(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, 'xrt_oe_surface_f64')
        return out

    def local_to_global(self, lb):
        """OE.local_to_global (oes/base.py:1165-1229) for a device Beam, in place: back-rotation, coherency matrix
        and amplitudes turned by roll + atan2(n_x, n_z), then virgin local -> global."""
        from .waves import _refuse_two_surface
        _refuse_two_surface(self, 'local_to_global')
        if self.dx:
            lb.x = lb.x + self.dx
        rots = []
        if self.extraPitch or self.extraRoll or self.extraYaw:
            rots += _rotation_list('-'+self.extraRotationSequence, self.extraPitch, self.extraRoll, self.extraYaw)
        rots += _rotation_list('-'+self.rotationSequence, self.pitch, self.roll + self.positionRoll, self.yaw)
        for ax, c, s in rots:
            if ax == 0:
                lb.y, lb.z = c*lb.y - s*lb.z, s*lb.y + c*lb.z
                lb.b, lb.c = c*lb.b - s*lb.c, s*lb.b + c*lb.c
            elif ax == 1:
                lb.x, lb.z = c*lb.x + s*lb.z, -s*lb.x + c*lb.z
                lb.a, lb.c = c*lb.a + s*lb.c, -s*lb.a + c*lb.c
            else:
                lb.x, lb.y = c*lb.x - s*lb.y, s*lb.x + c*lb.y
                lb.a, lb.b = c*lb.a - s*lb.b, s*lb.a + c*lb.b
        nx, ny, nz = self.local_n_device(lb.x, lb.y, lb.z)
        roll = (self.roll + self.positionRoll) + torch.atan2(nx, nz)
        c, s = torch.cos(roll), torch.sin(roll)
        c2, s2, cs = c**2, s**2, c * s
        Jss, Jpp, Jsp = lb.Jss, lb.Jpp, lb.Jsp
        JssN = Jss*c2 + Jpp*s2 + 2*Jsp.real*cs
        JppN = Jss*s2 + Jpp*c2 - 2*Jsp.real*cs
        JspN = torch.complex((Jpp-Jss)*cs + Jsp.real*(c2-s2), Jsp.imag.clone())
        lb.Jss, lb.Jpp, lb.Jsp = JssN, JppN, JspN
        if hasattr(lb, 'Es'):
            Es, Ep = lb.Es, lb.Ep
            lb.Es, lb.Ep = c*Es + s*Ep, -s*Es + c*Ep
        a0 = self.bl.sinAzimuth if self.bl is not None else 0.
        b0 = self.bl.cosAzimuth if self.bl is not None else 1.
        if a0 != 0:
            lb.a, lb.b = b0*lb.a - (-a0)*lb.b, (-a0)*lb.a + b0*lb.b
            lb.x, lb.y = b0*lb.x - (-a0)*lb.y, (-a0)*lb.x + b0*lb.y
        lb.x = lb.x + self.center[0]
        lb.y = lb.y + self.center[1]
        lb.z = lb.z + self.center[2]

    def _new_local(self, beam):
        lb = beam.like()
        lb.theta = torch.empty(len(beam), dtype=torch.float64, device=beam.device)
        if self.isParametric:
            for f in ('s', 'phi', 'r'):
                setattr(lb, f, torch.empty(len(beam), dtype=torch.float64, device=beam.device))
        return lb

    def reflect(self, beam=None, needLocal=True, noIntersectionSearch=False, returnLocalAbsorbed=None,
                checkTables=True, index_offset=0, allreduce=None, withFootprint=False, exposeOn=None):
        """Returns (beamGlobal, beamLocal) like OE.reflect (oes/reflect.py:17-163).

        *exposeOn* = a Screen: the kernel continues with Screen.expose (screens.py:225-302) and the method returns
        (screenBeam, beamLocal) -- bit for bit what `screen.expose(oe.reflect(beam)[0])` returns, without the global
        beam ever being written to (and read back from) HBM; for chains that only consume the screen."""
        pitch = self.pitch + (self.bragg if hasattr(self, 'bragg') else 0)
        params = self._make_params(pitch, self.roll + self.positionRoll, self.yaw, self.dx,
                                   createdByDiffract=bool(getattr(beam, 'createdByDiffract', False)))
        params.no_intersection = int(bool(noIntersectionSearch))  # reflect.py:676-681
        # the footprint / counter reductions are only paid for when somebody reads them
        params.want_footprint = int(bool(withFootprint or (checkTables and self.material is not None)))
        lb = self._new_local(beam) if needLocal else None
        if exposeOn is not None:
            gb, screen = None, exposeOn._fused_target(beam)
        else:
            gb, screen = beam.like(), None
        # an element with its own intersection method needs no bracket: pre-pass B is skipped
        phases = (1, 2) if (self.cu_find_intersection and not noIntersectionSearch) else None
        red = self._launch(params, self.material, beam, gb, lb, index_offset=index_offset, allreduce=allreduce,
                           phases=phases, screen=screen)
        if exposeOn is not None:
            gb = screen[1]
            gb.parentId = getattr(exposeOn, 'uuid', exposeOn.name)
        if checkTables and self.material is not None:
            self._check_tables(red)
        if withFootprint:
            self.footprint = [self._footprint_from(red)]
        lb_parent = getattr(self, 'uuid', self.name)
        if exposeOn is None:
            gb.parentId = lb_parent
        if lb is not None:
            lb.parentId = lb_parent
        return gb, (lb if needLocal else gb)


    def multiple_reflect(self, beam=None, maxReflections=1000, needElevationMap=False, returnLocalAbsorbed=None):
        """OE.multiple_reflect (oes/reflect.py:165-264): like reflect, with up to *maxReflections* reflections on the
        same surface.  Returns (beamGlobal, beamLocalN): beamGlobal carries *nRefl*; beamLocalN stacks the passes (all
        footprints, virgin-local coordinates as in the reference, where gb and lb are one beam), *nRefl* and *theta*
        included.  Every pass is the fused reflect kernel in its `multi` mode; passes after the first are preceded by
        the derivative stage (xrt_reflect_multi_deriv_f64: the ray leaves the surface before the next intersection is
        searched).  One host synchronisation per pass (the count of rays still alive ends the loop)."""
        if self.figureError is not None:
            # no reference behaviour to match: in the derivative stage find_dz does `surf += z_distorted` on the
            # NORMAL returned by local_n (base.py:823-830, :845) -- a tuple fails, a list is extended by the heights
            raise NotImplementedError('multiple_reflect with a figure error is outside the hot path')
        if self.material is not None and getattr(self.material, 'kind', 'mirror') not in ('auto', 'mirror', 'thin mirror'):
            raise NotImplementedError('multiple_reflect is on the hot path for mirrors (and bare surfaces)')
        lib = _lib.load()
        plugin = self._plugin()
        deriv = lib.xrt_reflect_multi_deriv_f64 if plugin is None else plugin.reflect_multi_deriv
        elevate = lib.xrt_reflect_multi_elevation_f64 if plugin is None else plugin.reflect_multi_elevation
        self.footprint = []
        n = len(beam)
        dev = beam.device
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        nRefl = torch.zeros(n, dtype=torch.int32, device=dev)
        tD = torch.empty(n, dtype=torch.float64, device=dev)
        cur = beam
        passes = []
        elev = None
        if needElevationMap:
            # elevationD = -1, elevationX / Y / Z = -maxHalfSizeOfOE until a ray has travelled between two impacts
            # (reflect.py:212-216); the derivative stage of every later pass renews the entries of its traced rays
            elev = [torch.full((n,), v, dtype=torch.float64, device=dev) for v in (-1., -1000., -1000., -1000.)]
        if int((beam.state > 0).sum()) == 0:  # reflect.py:200-201
            gb = Beam(copyFrom=beam)
            return gb, gb
        for iRefl in range(int(maxReflections)):
            params = self._make_params(self.pitch, self.roll + self.positionRoll, self.yaw, self.dx,
                                       from_global=(iRefl == 0), to_global=False)
            params.multi = 1
            params.good_mode = 0 if iRefl == 0 else 1  # state > 0, then state in (1, 2) (reflect.py:198, :228)
            out = cur.like()
            out.theta = torch.empty(n, dtype=torch.float64, device=dev)
            if iRefl == 0:
                red = self._launch(params, self.material, cur, out, None)
            else:
                # pre-pass A, derivative stage (scalars, roots), then pre-pass B + main pass from the roots
                red = self._launch(params, self.material, cur, out, None, phases=(1,))
                cin = cur.c_struct()
                for ph in (1, 2):
                    _lib.check(deriv(n, ctypes.byref(params), ctypes.byref(cin), red.data_ptr(), tD.data_ptr(), ph,
                                     stream), 'xrt_reflect_multi_deriv_f64')
                if elev is not None:  # find_dz at the derivative root (base.py:1285-1287; reflect.py:651-659)
                    _lib.check(elevate(n, ctypes.byref(params), ctypes.byref(cin), tD.data_ptr(),
                                       *[e.data_ptr() for e in elev], stream), 'xrt_reflect_multi_elevation_f64')
                params.multi_tmin = tD.data_ptr()
                red = self._launch(params, self.material, cur, out, None, phases=(6,))
            self._check_tables(red)
            self.footprint.append(self._footprint_from(red))
            good = (out.state == 1) | (out.state == 2)
            nRefl = nRefl + good.to(torch.int32)
            out.nRefl = nRefl.clone()
            if elev is not None:
                for k, e in zip(('elevationD', 'elevationX', 'elevationY', 'elevationZ'), elev):
                    setattr(out, k, e.clone())
            ngood, ngoodN = torch.stack((good.sum(), (out.state == 1).sum())).tolist()
            if ngoodN == 0 and passes:
                out.theta = passes[-1].theta.clone()  # lb.theta is only renewed when a ray reflects (reflect.py:789)
            passes.append(out)
            cur = out
            if ngood == 0:
                break
        lbN = passes[0]
        for b in passes[1:]:
            lbN.concatenate(b)
        # in global coordinates (reflect.py:245-253)
        gb = Beam(copyFrom=cur)
        gb.nRefl = nRefl
        goodAfter = nRefl > 0
        gb.state = torch.where(goodAfter, torch.ones_like(gb.state), gb.state)
        a0 = self.bl.sinAzimuth if self.bl is not None else 0.
        b0 = self.bl.cosAzimuth if self.bl is not None else 1.
        ga, gbb, gx, gy = gb.a, gb.b, gb.x, gb.y
        if a0 != 0:  # rotate_z(., ., b0, -a0)
            ga, gbb = b0*gb.a - (-a0)*gb.b, (-a0)*gb.a + b0*gb.b
            gx, gy = b0*gb.x - (-a0)*gb.y, (-a0)*gb.x + b0*gb.y
        gx = gx + self.center[0]
        gy = gy + self.center[1]
        gz = gb.z + self.center[2]
        # not intersected rays remain unchanged except their state (copy_beam without the state, :251-253)
        for f, v in (('x', gx), ('y', gy), ('z', gz), ('a', ga), ('b', gbb), ('c', gb.c)):
            setattr(gb, f, torch.where(goodAfter, v, getattr(beam, f)))
        for f in ('path', 'E', 'Jss', 'Jpp', 'Jsp', 'Es', 'Ep'):
            if hasattr(gb, f):
                setattr(gb, f, torch.where(goodAfter, getattr(gb, f), getattr(beam, f)))
        if hasattr(gb, 'theta'):
            del gb.theta
        gb.parentId = lbN.parentId = getattr(self, 'uuid', self.name)
        return gb, lbN

    # ------------------------------------------------------------------ wave mode (oes/reflect.py:266-449)
    def _surface_centre_global(self):
        """The middle point of limPhysX x limPhysY on the surface, in global coordinates (reflect.py:344-361)."""
        cx = (self.limPhysX[1] + self.limPhysX[0])*0.5
        cy = (self.limPhysY[1] + self.limPhysY[0])*0.5
        dev = torch.device('cuda', torch.cuda.current_device())
        lbc = Beam(1, device=dev)
        lbc.x.fill_(cx)
        lbc.y.fill_(cy)
        lbc.z = self.local_z_device(lbc.x, lbc.y)
        self.local_to_global(lbc)
        return [float(lbc.x[0]), float(lbc.y[0]), float(lbc.z[0])]

    def prepare_wave(self, prevOE, nrays, shape='auto', area='auto', rw=None, xy=None, device=None):
        """OE.prepare_wave (oes/reflect.py:266-403): samples on this surface as receiving points for
        waves.diffract.  *nrays*: int (random samples; *xy* optionally supplies the (nrays, 2) uniform deviates
        the reference draws with np.random.rand), (nx, ny) mesh sizes, or (xarray, yarray)."""
        if rw is None:
            from . import waves as rw
        dev = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
        if isinstance(nrays, (int, float)):
            nsamples = int(nrays)
            if xy is None:
                xy = np.random.rand(nsamples, 2)
            if shape == 'auto':
                shape = self.shape
            if not shape.startswith('re'):
                raise NotImplementedError("only shape='rect' is on the hot path")
            dX = self.limPhysX[1] - self.limPhysX[0]
            dY = self.limPhysY[1] - self.limPhysY[0]
            x = xy[:, 0] * dX + self.limPhysX[0]
            y = xy[:, 1] * dY + self.limPhysY[0]
            if area == 'auto':
                area = dX * dY
        elif isinstance(nrays, (list, tuple)):
            if isinstance(nrays[0], (int, float)):
                xx = np.linspace(*self.limPhysX, nrays[0])
                yy = np.linspace(*self.limPhysY, nrays[1])
            elif isinstance(nrays[0], np.ndarray) and isinstance(nrays[1], np.ndarray):
                xx, yy = nrays
            else:
                raise ValueError('wrong type of `nrays`!')
            X, Y = np.meshgrid(xx, yy)
            x = X.ravel()
            y = Y.ravel()
            nsamples = len(x)
            if area == 'auto':
                dX = self.limPhysX[1] - self.limPhysX[0]
                dY = self.limPhysY[1] - self.limPhysY[0]
                area = dX * dY
        else:
            raise ValueError('wrong type of `nrays`!')
        lb = Beam(nsamples, withAmplitudes=True, device=dev)
        lb.state.fill_(1)
        lb.parentId = getattr(prevOE, 'uuid', getattr(prevOE, 'name', None))
        lb.x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(dev)
        lb.y = torch.from_numpy(np.ascontiguousarray(y, dtype=np.float64)).to(dev)
        lb.z = self.local_z_device(lb.x, lb.y)
        self.local_to_global(lb)
        if hasattr(prevOE, 'rotationSequence'):  # OE: the middle of its surface, not its centre
            prevCenter = prevOE._surface_centre_global()
        else:
            prevCenter = [float(v) for v in prevOE.center]
        lb.a = lb.x - prevCenter[0]
        lb.b = lb.y - prevCenter[1]
        lb.c = lb.z - prevCenter[2]
        norm = (lb.a**2 + lb.b**2 + lb.c**2)**0.5
        lb.a = lb.a / norm
        lb.b = lb.b / norm
        lb.c = lb.c / norm
        lb.x = torch.full_like(lb.a, prevCenter[0])
        lb.y = torch.full_like(lb.a, prevCenter[1])
        lb.z = torch.full_like(lb.a, prevCenter[2])
        # the unit z vector of this OE seen from prevCenter (reflect.py:376-385)
        lbn = Beam(1, device=dev)
        lbn.b.fill_(0.)
        lbn.c.fill_(1.)
        self.local_to_global(lbn)
        nh = lbn.to_numpy()
        a = nh['x'] - prevCenter[0]
        b = nh['y'] - prevCenter[1]
        c = nh['z'] - prevCenter[2]
        norm = (a**2 + b**2 + c**2)**0.5
        areaNormalFact = abs(float(((a*nh['a'][0] + b*nh['b'][0] + c*nh['c'][0]) / norm)[0]))

        waveGlobal, waveLocal = self.reflect(lb)
        good = (waveLocal.state == 1) | (waveLocal.state == 2)
        ngood = int(good.sum())
        waveGlobal.filter_by_index(good)
        waveLocal.filter_by_index(good)
        area *= ngood / float(nsamples)
        waveLocal.area = area
        waveLocal.areaNormal = area * areaNormalFact
        waveLocal.dS = area / float(nsamples)
        waveLocal.toOE = self
        waveLocal.parentId = getattr(self, 'uuid', self.name)
        rw.prepare_wave(prevOE, waveLocal, waveGlobal.x, waveGlobal.y, waveGlobal.z)
        return waveLocal

    def propagate_wave(self, wave=None, beam=None, nrays='auto', prevOE=None, xy=None):
        """OE.propagate_wave (oes/reflect.py:405-449): Kirchhoff-propagates *wave* (the local beam on the
        previous element) onto this surface and reflects it.  Returns (beamGlobal, beamLocal).  The previous
        element is looked up through bl.oesDict like in xrt, or given as *prevOE*; auto-alignment is xrt's
        host-side business and is not repeated here."""
        from . import waves as rw
        waveSize = len(wave) if nrays == 'auto' else int(nrays)
        if prevOE is None:
            prevOE = getattr(wave, 'toOE', None) or self.bl.oesDict[wave.parentId][0]
        waveOnSelf = self.prepare_wave(prevOE, waveSize, rw=rw, xy=xy, device=wave.device)
        if hasattr(prevOE, 'shine'):
            beamToSelf = prevOE.shine(wave=waveOnSelf)
            nIS = False
        else:
            beamToSelf = rw.diffract(wave, waveOnSelf)
            nIS = True
        retGlo, retLoc = self.reflect(beamToSelf, noIntersectionSearch=nIS)
        retLoc.parentId = getattr(self, 'uuid', self.name)
        return retGlo, retLoc


class ToroidMirror(OE):
    """z = y^2/2R + r (1 - sqrt(1 - x^2/r^2))  (oes/__init__.py:320-411)"""
    surfaceKind = 1

    def __init__(self, *args, **kwargs):
        R = kwargs.pop('R', 5.0e6)
        r = kwargs.pop('r', 50.)
        OE.__init__(self, *args, **kwargs)
        self.R = 1e100 if R in (0, None) else R
        self.r = 1e100 if r in (0, None) else r
        if isinstance(self.R, (list, tuple)) or isinstance(self.r, (list, tuple)):
            raise NotImplementedError('give R and r as numbers (Coddington (p, q) tuples are resolved by xrt)')

    def _surface_params(self):
        return [self.R, self.r]


def _abs_pitch_to_axis(oe, axis=(0, 1, 0)):
    """|angle| between the mirror surface and *axis* (global): the mirror normal (0, 0, 1) taken to global coordinates
    as in EllipticalMirrorParam._reset_pq / ParabolicalMirrorParam._reset_pq (parametric.py:84-96, :113-117, :379-396)."""
    n = np.array([0., 0., 1.])
    for ax, c, s in _rotation_list('-'+oe.rotationSequence, oe.pitch, oe.roll+oe.positionRoll, oe.yaw):
        i, j = {0: (1, 2), 1: (0, 2), 2: (0, 1)}[ax]
        if ax == 1:
            n[i], n[j] = c*n[i] + s*n[j], -s*n[i] + c*n[j]
        else:
            n[i], n[j] = c*n[i] - s*n[j], s*n[i] + c*n[j]
    a0 = oe.bl.sinAzimuth if oe.bl is not None else 0.
    b0 = oe.bl.cosAzimuth if oe.bl is not None else 1.
    if a0 != 0:
        n[0], n[1] = b0*n[0] - (-a0)*n[1], (-a0)*n[0] + b0*n[1]
    norm = sum([a**2 for a in axis])**0.5
    sintheta = sum([a*nn for a, nn in zip(axis, n)]) / norm
    return abs(np.arcsin(sintheta))


SimpleVFM = ToroidMirror  # oes/__init__.py:414


class NormalFZP(OE):
    """Circular Fresnel zone plate in the local xy plane, optical axis along local z (oes/gratings.py:10-137): zones of
    zero thickness with transmittivity 0 and 1.  Give the material kind='FZP' (or leave it 'auto').  The zone mask
    and the radial groove vector are evaluated per ray inside the fused kernel from the zone radii *rn*."""

    def __init__(self, *args, **kwargs):
        self.f = kwargs.pop('f', 50)
        self.E = kwargs.pop('E', 1000)
        self.N = kwargs.pop('N', 1000)
        self.isCentralZoneBlack = kwargs.pop('isCentralZoneBlack', True)
        self.thinnestZone = kwargs.pop('thinnestZone', None)
        self.reset()
        kwargs['limPhysX'] = [-self.rn[-1], self.rn[-1]]
        kwargs['limPhysY'] = [-self.rn[-1], self.rn[-1]]
        OE.__init__(self, *args, **kwargs)
        if self.material is None or getattr(self.material, 'kind', 'auto') not in ('FZP', 'auto'):
            raise ValueError("NormalFZP needs a material of kind='FZP'")

    def reset(self):  # gratings.py:94-117
        from .physconsts import CH
        lambdaE = CH / self.E * 1e-7
        if self.thinnestZone is not None:
            self.N = lambdaE * self.f / 4. / self.thinnestZone**2
        self.zones = np.arange(self.N+1)
        self.rn = np.sqrt(self.zones*self.f*lambdaE + 0.25*(self.zones*lambdaE)**2)
        self._rn_dev = (None,)


class EllipticalMirrorParam(OE):
    """Elliptical cylinder / ellipsoid in the parametric (s, phi, r) representation
    (oes/parametric.py:9-249).  The ellipse parameters follow _reset_pq (:104-150) for the default p-arm
    along the global y axis."""
    surfaceKind = 2
    isParametric = True

    def __init__(self, *args, **kwargs):
        self.p = kwargs.pop('p', 1000)
        self.q = kwargs.pop('q', 1000)
        self.isCylindrical = kwargs.pop('isCylindrical', False)
        if kwargs.pop('isClosed', False):
            raise NotImplementedError('isClosed=True is outside the hot path')
        for k in ('f1', 'f2', 'pAxis'):
            if kwargs.pop(k, None) is not None:
                raise NotImplementedError('{0} is resolved by xrt; give p and q'.format(k))
        OE.__init__(self, *args, **kwargs)
        self._reset_pq()

    def _reset_pq(self):
        absPitch = _abs_pitch_to_axis(self)
        gamma = np.arctan2((self.p - self.q) * np.sin(absPitch), (self.p + self.q) * np.cos(absPitch))
        self.cosGamma = np.cos(gamma)
        self.sinGamma = np.sin(gamma)
        self.y0 = (self.q - self.p)/2. * np.cos(absPitch)
        self.z0 = (self.q + self.p)/2. * np.sin(absPitch)
        self.ellipseA = (self.q + self.p)/2.
        self.ellipseB = np.sqrt(self.q * self.p) * np.sin(absPitch)

    def _surface_params(self):
        return [self.ellipseA, self.ellipseB, self.y0, self.z0, self.cosGamma, self.sinGamma,
                1. if self.isCylindrical else 0.]


class DCM(OE):
    """Double-crystal monochromator: two flat crystals (oes/dcm.py:12-354)."""

    def __init__(self, *args, **kwargs):
        self.bragg = kwargs.pop('bragg', 0)
        self.cryst1roll = kwargs.pop('cryst1roll', 0)
        self.cryst2roll = kwargs.pop('cryst2roll', 0)
        self.cryst2pitch = kwargs.pop('cryst2pitch', 0)
        self.cryst2finePitch = kwargs.pop('cryst2finePitch', 0)
        self.cryst2perpTransl = kwargs.pop('cryst2perpTransl', 0)
        self.cryst2longTransl = kwargs.pop('cryst2longTransl', 0)
        self.limPhysX2 = kwargs.pop('limPhysX2', [-maxHalfSizeOfOE, maxHalfSizeOfOE])
        self.limPhysY2 = kwargs.pop('limPhysY2', [-maxHalfSizeOfOE, maxHalfSizeOfOE])
        self.limOptX2 = kwargs.pop('limOptX2', None)
        self.limOptY2 = kwargs.pop('limOptY2', None)
        self.material2 = kwargs.pop('material2', None)
        fixedOffset = kwargs.pop('fixedOffset', None)
        OE.__init__(self, *args, **kwargs)
        if self.material2 is None:
            self.material2 = self.material
        if isinstance(self.bragg, str):
            raise NotImplementedError("bragg='auto'/energy strings are resolved by xrt's auto_align; give radians")
        self.fixedOffset = fixedOffset
        if fixedOffset not in [0, None]:
            self.cryst2perpTransl = fixedOffset/2./np.cos(self.bragg)  # dcm.py:240-245

    def reflect(self, *a, **k):
        raise AttributeError('use DCM.double_reflect')

    def double_reflect(self, beam=None, needLocal=True, fromVacuum1=True, fromVacuum2=True,
                       returnLocalAbsorbed=None, checkTables=True, index_offset=0, allreduce=None, exposeOn=None):
        """Returns (beamGlobal, beamLocal1, beamLocal2) like DCM.double_reflect (oes/dcm.py:247-354).  *exposeOn*: see
        OE.reflect -- the second crystal's kernel continues with Screen.expose, beamGlobal is the screen beam."""
        p1 = self._make_params(self.pitch + self.bragg, self.roll + self.positionRoll + self.cryst1roll,
                               self.yaw, self.dx, to_global=False)
        p1.from_vacuum = int(bool(fromVacuum1))
        gb = beam.like()
        lo1 = self._new_local(beam) if needLocal else None
        red = self._launch(p1, self.material, beam, gb, lo1, index_offset=index_offset, allreduce=allreduce)
        if checkTables and self.material is not None:
            self._check_tables(red)
        p2 = self._make_params(-self.pitch - self.bragg + self.cryst2pitch + self.cryst2finePitch,
                               self.roll + self.cryst2roll + self.positionRoll, -self.yaw,
                               -self.dx, self.cryst2longTransl, -self.cryst2perpTransl, is2ndXtal=True,
                               from_global=False, to_global=True)
        p2.good_mode = 1
        p2.zero_dead_local = 1
        p2.from_vacuum = int(bool(fromVacuum2))
        p2.dead_state_lost = int(hasattr(self, 't'))  # is instance of Plate (dcm.py:311-312, :331-332)
        lo2 = self._new_local(beam) if needLocal else None
        if exposeOn is not None and not hasattr(self, 't') and self._plugin() is None:
            gb2, screen = None, exposeOn._fused_target(beam)
        else:
            gb2, screen = beam.like(), None
        red = self._launch(p2, self.material2, gb, gb2, lo2, fallback=beam, index_offset=index_offset,
                           allreduce=allreduce, screen=screen)
        if screen is not None:
            gb2 = screen[1]
            gb2.parentId = getattr(exposeOn, 'uuid', exposeOn.name)
        elif exposeOn is not None:
            gb2 = exposeOn.expose(gb2)
        if checkTables and self.material2 is not None:
            self._check_tables(red)
        if not needLocal:
            lo1 = lo2 = gb2
        return gb2, lo1, lo2


class Plate(DCM):
    """A body with two surfaces (oes/refractive.py:12-234): derived from DCM like the reference's, the second
    'crystal' is the exit surface at cryst2perpTransl = -t, tilted by wedgeAngle."""

    def __init__(self, *args, **kwargs):
        self.t = kwargs.pop('t', 0)
        self.wedgeAngle = kwargs.pop('wedgeAngle', 0)
        kwargs.setdefault('overEdge', '')
        DCM.__init__(self, *args, **kwargs)
        self.cryst2perpTransl = -self.t
        self.cryst2pitch = self.wedgeAngle
        # refractive.py:39-48 mirrors the x limits only `if isinstance(self.limPhysX, (list, tuple))`; by then the
        # property setter has wrapped them in raycing.Limits, so the reference ends up with limits 2 == limits 1
        # (verified on its outputs: tests/golden/refract_*.npz carry surfPhysX2 == surfPhysX).
        self.limPhysX2 = self.limPhysX
        self.limPhysY2 = self.limPhysY
        self.limOptX2 = self.limOptX
        self.limOptY2 = self.limOptY
        self.material2 = self.material

    def double_reflect(self, *a, **k):
        raise AttributeError('use Plate.double_refract')

    def double_refract(self, beam=None, needLocal=True, returnLocalAbsorbed=None, **kw):
        """Returns (beamGlobal, beamLocal1, beamLocal2) like Plate.double_refract (oes/refractive.py:171-234)."""
        return DCM.double_reflect(self, beam=beam, needLocal=needLocal, fromVacuum1=True, fromVacuum2=False, **kw)


class ParaboloidFlatLens(Plate):
    """Refractive lens / CRL stack, z = (x^2 + y^2) / (4 focus) limited to zmax (oes/refractive.py:237-549).  As
    in the reference, local_z2 resolves to local_z -> local_z1 (:401-403, :424-425): both interfaces carry the
    parabolic profile."""
    surfaceKind = 3
    isCylinder = False
    isDouble = False

    def __init__(self, *args, **kwargs):
        focus = kwargs.pop('focus', 1.)
        self.zmax = kwargs.pop('zmax', None)
        nCRL = kwargs.pop('nCRL', 1)
        kwargs.setdefault('pitch', np.pi/2)
        Plate.__init__(self, *args, **kwargs)
        if isinstance(focus, (list, tuple)) or isinstance(nCRL, (list, tuple)):
            raise NotImplementedError('give focus and nCRL as numbers (the (f, E) forms need the host refractive '
                                      'index: use get_nCRL / get_focus)')
        self.focus = focus
        self.nCRL = max(int(round(nCRL)), 1)  # refractive.py:340-344

    def _surface_params(self):
        return [self.focus, self.zmax if self.zmax is not None else 0., 0. if self.zmax is None else 1.,
                1. if self.isCylinder else 0.]

    def get_nCRL(self, f, E):  # refractive.py:430-441
        nFactor = 0.5 if self.isDouble else 1.
        df = 1. - self.material.get_refractive_index(E).real
        return self.focus / (f*df) * (2*nFactor)

    def get_focus(self, f, E):  # refractive.py:443-455
        nFactor = 0.5 if self.isDouble else 1.
        df = 1. - self.material.get_refractive_index(E).real
        return (f*df) * self.nCRL / (2*nFactor)

    def multiple_refract(self, beam=None, needLocal=True, returnLocalAbsorbed=None, **kw):
        """Sequentially applies double_refract to the nCRL lenslets, each shifted by zmax + t (2 zmax + t for the
        double lenses) along the local [0, 0, -1] (oes/refractive.py:457-549).  Returns (beamGlobal of the last
        lens, beamLocal1, beamLocal2 of the first one)."""
        if self.nCRL == 1:
            self.centerShift = np.zeros(3)
            return self.double_refract(beam=beam, needLocal=needLocal, **kw)
        tempCenter = [c for c in self.center]
        beamIn = beam
        zmax = 5 if self.zmax is None else self.zmax
        step = 2.*zmax + self.t if self.isDouble else zmax + self.t
        toward = [0, -step, 0]
        for ilens in range(self.nCRL):
            lglobal, tlocal1, tlocal2 = self.double_refract(beam=beamIn, needLocal=needLocal, **kw)
            if self.zmax is not None:
                toward = _rotate_point([0, 0, 1], self.rotationSequence, self.pitch, self.roll+self.positionRoll,
                                       self.yaw)
                self.center[0] -= step * toward[0]
                self.center[1] -= step * toward[1]
                self.center[2] -= step * toward[2]
            beamIn = lglobal
            if ilens == 0:
                llocal1, llocal2 = tlocal1, tlocal2
        self.centerShift = step * np.array(toward)
        self.center = tempCenter
        return lglobal, llocal1, llocal2


class ParabolicCylinderFlatLens(ParaboloidFlatLens):
    """oes/refractive.py:552-617: the profile depends on y only."""
    isCylinder = True


class DoubleParaboloidLens(ParaboloidFlatLens):
    """oes/refractive.py:620-652 (stack step 2 zmax + t)."""
    isDouble = True


class DoubleParabolicCylinderLens(ParabolicCylinderFlatLens):
    """oes/refractive.py:655-663."""
    isDouble = True


def _rotate_point(point, rotationSequence='RzRyRx', pitch=0, roll=0, yaw=0):
    """raycing.rotate_point (_rotate.py:87-108)."""
    newp = [coord for coord in point]
    ind1 = {0: 1, 1: 0, 2: 0}
    ind2 = {0: 2, 1: 2, 2: 1}
    for ax, cA, sA in _rotation_list(rotationSequence, pitch, roll, yaw):
        u, v = newp[ind1[ax]], newp[ind2[ax]]
        if ax == 1:  # rotate_y
            newp[ind1[ax]], newp[ind2[ax]] = cA*u + sA*v, -sA*u + cA*v
        else:  # rotate_x, rotate_z
            newp[ind1[ax]], newp[ind2[ax]] = cA*u - sA*v, sA*u + cA*v
    return newp


_SPLICED = ('BentFlatMirror', 'SimpleVCM', 'ConicalMirror', 'JohannCylinder', 'JohanssonCylinder', 'JohannToroid',
            'ParabolicalMirrorParam', 'ParabolicMirror', 'LauePlate', 'BentLaueCylinder', 'BlazedGrating')


def __getattr__(name):
    """The curved elements whose surfaces are spliced CUDA texts (xrt_b200/oes_spliced.py) carry xrt's names in this
    namespace too (resolved lazily: that module derives its classes from OE above)."""
    if name in _SPLICED:
        from . import oes_spliced
        return getattr(oes_spliced, name)
    raise AttributeError(name)
