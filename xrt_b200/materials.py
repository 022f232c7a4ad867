# -*- coding: utf-8 -*-
"""Materials of the hot path: mirror coatings (Fresnel amplitude) and perfect fcc/diamond crystals (thick
Bragg amplitude).  Host side = parameters + tables; the per-ray arithmetic runs in the CUDA kernels
(xrt_b200/csrc/reflect.cu) inside OE.reflect, or stand-alone through `get_amplitude`.

Mirrors the constructors of xrt's Material (xrt/backends/raycing/materials/material.py:26) and
CrystalFcc / CrystalDiamond / CrystalSi (materials/crystals_basic.py:8-115): same keywords, same defaults
(table 'Chantler total'), same derived quantities (mass, d, V, chiToF).  Scattering-factor tables:
xrt_b200/data/elements.npz (see tools/extract_tables.py).
"""
import ctypes
import os

import numpy as np

from .physconsts import PI, R0, CH, AVOGADRO, PI2

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'elements.npz')
_cache = None

XRT_MAX_ELEM = 4


class xrt_material_params(ctypes.Structure):
    """struct xrt_material_params of include/xrt_b200.h"""
    _fields_ = [('kind', ctypes.c_int), ('nelem', ctypes.c_int),
                ('Z', ctypes.c_double * XRT_MAX_ELEM), ('quantity', ctypes.c_double * XRT_MAX_ELEM),
                ('tab_off', ctypes.c_int * XRT_MAX_ELEM), ('tab_len', ctypes.c_int * XRT_MAX_ELEM),
                ('rho', ctypes.c_double), ('mass', ctypes.c_double), ('t', ctypes.c_double),
                ('d', ctypes.c_double), ('chiToF', ctypes.c_double), ('factDW', ctypes.c_double),
                ('f0c', ctypes.c_double * 11), ('d2f_re', ctypes.c_double), ('d2f_im', ctypes.c_double),
                ('f0scale', ctypes.c_double), ('hkl_allowed', ctypes.c_int), ('geom', ctypes.c_int),
                ('multilayer', ctypes.c_void_p)]


class xrt_multilayer_params(ctypes.Structure):
    """struct xrt_multilayer_params of include/xrt_b200.h"""
    _fields_ = [('tLayer', xrt_material_params), ('bLayer', xrt_material_params), ('substrate', xrt_material_params),
                ('nPairs', ctypes.c_int), ('uniform', ctypes.c_int), ('idThickness', ctypes.c_double),
                ('d', ctypes.c_double), ('dti', ctypes.c_void_p), ('dbi', ctypes.c_void_p),
                ('substRoughness', ctypes.c_double), ('transmitted', ctypes.c_int),
                ('substThickness', ctypes.c_double), ('specular', ctypes.c_int)]


def _tables():
    global _cache
    if _cache is None:
        with np.load(_DATA) as f:
            _cache = {k: np.array(f[k]) for k in f.files}
    return _cache


class Element(object):
    """f0, f1, f2 of one chemical element (materials/element.py:92-263)."""

    def __init__(self, elem=None, table='Chantler'):
        t = _tables()
        if elem + '_Z' not in t:
            raise ValueError('element {0!r} is not in xrt_b200/data/elements.npz; add it to '
                             'tools/extract_tables.py'.format(elem))
        if not table.startswith('Chantler'):
            raise ValueError("only the 'Chantler' / 'Chantler total' tables are packaged")
        self.name = elem
        self.table = table
        self.Z = int(t[elem + '_Z'])
        self.mass = float(t[elem + '_mass'])
        self.f0coeffs = [float(v) for v in t[elem + '_f0']]
        f2key = '_f2tot' if 'total' in table else '_f2'  # element.py:212
        # float32 in the data files; np.interp promotes to float64 (element.py:261-262)
        self.E = t[elem + '_E'].astype(np.float64)
        self.f1 = t[elem + '_f1'].astype(np.float64)
        self.f2 = t[elem + f2key].astype(np.float64)

    def get_f0(self, qOver4pi=0):
        return self.f0coeffs[5] + sum(a * np.exp(-b * qOver4pi**2)
                                      for a, b in zip(self.f0coeffs[:5], self.f0coeffs[6:]))


class EmptyMaterial(object):
    """materials/__init__.py:101-113: a material whose only property is its kind (a grating without a coating):
    no get_amplitude, i.e. reflectivity 1."""

    def __init__(self, kind='grating'):
        self.kind = kind
        self.geom = ''
        self.name = ''
        self.efficiency = None


class Material(object):
    """Mirror coating / bulk material; kinds 'mirror', 'thin mirror' and 'grating' are on the hot path."""

    def __init__(self, elements=None, quantities=None, kind='auto', rho=0, t=None, table='Chantler total',
                 efficiency=None, name=''):
        self.efficiency = efficiency  # [(order, value), ...] for kind 'grating' (material.py:78-93)
        if isinstance(elements, str):
            elements = elements,
        self.table = table
        self.elements = [Element(e, table) for e in elements]
        if quantities is None:
            self.quantities = [1. for e in self.elements]
        elif not isinstance(quantities, (list, tuple)):
            self.quantities = [quantities]
        else:
            self.quantities = list(quantities)
        if len(self.elements) > XRT_MAX_ELEM:
            raise ValueError('at most {0} elements per material'.format(XRT_MAX_ELEM))
        self.kind = kind
        self.rho = rho
        self.t = t
        self.name = name or ''.join(e.name for e in self.elements)
        self.mass = 0.
        for elem, xi in zip(self.elements, self.quantities):  # material.py:262-275
            self.mass += xi * elem.mass
        self._dev_tables = {}

    # ------------------------------------------------------------------ device plumbing
    def _kind_code(self):
        kind = 'mirror' if self.kind == 'auto' else self.kind  # material.py:450-451
        if kind in ('mirror', 'grating'):  # a grating reflects like a mirror (material.py:475)
            return 1
        if kind == 'thin mirror':
            if self.t is None:
                raise ValueError('Give the thin mirror a thickness!')
            return 2
        if kind in ('plate', 'lens'):  # transmittivity + refraction (material.py:482-486, reflect.py:894-919)
            return 5
        raise NotImplementedError("material kind {0!r} is outside the hot path (SURVEY.md section 8f)".format(kind))

    def host_tables(self):
        """Concatenated [E | f1 | f2] tables of all elements and their slices."""
        offs, lens, E, f1, f2 = [], [], [], [], []
        pos = 0
        for e in self.elements:
            offs.append(pos)
            lens.append(len(e.E))
            pos += len(e.E)
            E.append(e.E)
            f1.append(e.f1)
            f2.append(e.f2)
        tab = np.concatenate([np.concatenate(E), np.concatenate(f1), np.concatenate(f2)])
        return tab, pos, offs, lens

    def device_tables(self, device):
        import torch
        key = str(device)
        if key not in self._dev_tables:
            tab, n, offs, lens = self.host_tables()
            self._dev_tables[key] = (torch.from_numpy(tab).to(device), n, offs, lens)
        return self._dev_tables[key]

    def c_struct(self, offs, lens):
        m = xrt_material_params()
        m.kind = self._kind_code()
        m.nelem = len(self.elements)
        for i, (e, xi) in enumerate(zip(self.elements, self.quantities)):
            m.Z[i] = float(e.Z)
            m.quantity[i] = float(xi)
            m.tab_off[i] = offs[i]
            m.tab_len[i] = lens[i]
        m.rho = float(self.rho)
        m.mass = float(self.mass)
        m.t = float(self.t) if self.t is not None else 0.
        return m

    def get_refractive_index(self, E):
        """Host (numpy) refractive index, material.py:348-378 -- used by alignment helpers only (get_dtheta of
        multilayers); the per-ray evaluation lives in the CUDA kernels."""
        E = np.asarray(E, dtype=float)
        xf = np.zeros_like(E) * 0j
        for elem, xi in zip(self.elements, self.quantities):
            if np.any(E < elem.E[0]) or np.any(E > elem.E[-1]):
                raise ValueError('E is out of the f1/f2 data table range')
            f1 = np.interp(E, elem.E, elem.f1)
            f2 = np.interp(E, elem.E, elem.f2)
            xf += (elem.Z + f1 + 1j*f2) * xi
        return 1 - 1e-24 * AVOGADRO * R0 / PI2 * (CH/E)**2 * self.rho * xf / self.mass

    # ------------------------------------------------------------------ stand-alone amplitude (a12)
    def get_amplitude(self, E, beamInDotNormal, fromVacuum=True):
        """(rs, rp) on host arrays through the CUDA kernel (material.py:415-493, kinds mirror / thin mirror).
        The absorption outputs mu and n.real*k of the reference tuple are not on the hot path."""
        if not fromVacuum:
            raise NotImplementedError('fromVacuum=False (refraction) is outside the hot path')
        from . import kernels
        return kernels.material_amplitude_host(self, E, beamInDotNormal)


class CrystalFcc(Material):
    """Perfect one-element fcc crystal (materials/crystals_basic.py:8-31 with the Crystal constructor
    crystal.py:39-181): thick Bragg case, or thickness *t* (mm) in Bragg / Laue geometry, reflected / transmitted."""
    structure = 'fcc'

    def __init__(self, hkl=(1, 1, 1), d=0, V=None, elements='Si', quantities=None, rho=0, t=None, factDW=1.,
                 geom='Bragg reflected', table='Chantler total', name='', **kwargs):
        super().__init__(elements, quantities, kind='crystal', rho=rho, t=t, table=table, name=name)
        if len(self.elements) != 1:
            raise NotImplementedError('multi-atom structure factors are outside the hot path')
        if not (geom.startswith('Bragg') or geom.startswith('Laue')) or geom.endswith('Fresnel'):
            raise NotImplementedError('geom {0!r}: Bragg / Laue, reflected / transmitted are on this path'.format(geom))
        if t is None and geom != 'Bragg reflected':
            raise ValueError('Laue and transmitted geometries need a crystal thickness t')
        self.hkl = tuple(int(i) for i in hkl)
        self.sqrthkl2 = (sum(i**2 for i in self.hkl))**0.5
        self.d = d
        self.V = V if V is not None else (d * self.sqrthkl2)**3  # crystal.py:172-176
        self.chiToF = -R0 / PI / self.V  # minus! (crystal.py:179)
        self.geom = geom
        self.factDW = factDW

    def _kind_code(self):
        return 3

    def get_Bragg_angle(self, E, order=1):
        a = np.array(order * CH / (2*self.d*E))
        a[a > 1] = 1 - 1e-16
        a[a < -1] = -1 + 1e-16
        return np.arcsin(a)

    def _diamond_to_fcc(self):
        return 1. + 0j

    # alignment helpers on the host (numpy): per-element scalars a beamline script computes once, not per-ray work
    def get_dtheta_symmetric_Bragg(self, E):
        """crystal.py:1122-1139: chi0 / sin(2 thetaB), F0 from the tables (crystals_basic.py:22-31, :76-80)."""
        E = np.asarray(E, dtype=float)
        elem = self.elements[0]
        if np.any(E < elem.E[0]) or np.any(E > elem.E[-1]):
            raise ValueError('E is out of the f1/f2 data table range')
        anomalous = np.interp(E, elem.E, elem.f1) + 1j*np.interp(E, elem.E, elem.f2)
        F0 = 4 * (elem.Z + anomalous) * self.factDW
        if self.structure == 'diamond':
            F0 = F0 * 2
        chi0 = F0 * (self.chiToF * (CH / E)**2)
        return (chi0 / np.sin(2*self.get_Bragg_angle(E))).real

    def get_dtheta(self, E, alpha=None):
        """Crystal.get_dtheta (crystal.py:1141-1173), Bragg geometry: the refraction correction to the Bragg angle
        a script subtracts when it sets `bragg` (what auto_align does, beamline.py:636-681)."""
        if alpha is None:
            alpha = 0
        thetaB = self.get_Bragg_angle(E)
        pm = -1
        gamma0 = np.sin(thetaB + alpha)
        gammah = pm * np.sin(thetaB - alpha)
        symm_dt = self.get_dtheta_symmetric_Bragg(E)
        osqg0 = np.sqrt(1. - gamma0**2)
        dtheta0 = (pm*gamma0 - pm*np.sqrt(gamma0**2 + pm*(gamma0 - gammah) * osqg0 * symm_dt)) / osqg0
        return -dtheta0

    def c_struct(self, offs, lens):
        m = super().c_struct(offs, lens)
        m.d = float(self.d)
        m.chiToF = float(self.chiToF)
        m.factDW = float(self.factDW)
        residue = sum(i % 2 for i in self.hkl)
        m.hkl_allowed = 1 if residue in (0, 3) else 0
        m.geom = (1 if self.geom.startswith('Laue') else 0) | (2 if self.geom.endswith('transmitted') else 0)
        f0 = self.elements[0].get_f0(0.5 / self.d)  # crystal.py:630 get_F_chi(E, 0.5/crystd)
        m.f0c[0] = float(f0)
        d2f = self._diamond_to_fcc()
        m.d2f_re, m.d2f_im = float(d2f.real), float(d2f.imag)
        m.f0scale = 2. if self.structure == 'diamond' else 1.
        return m

    def get_amplitude(self, E, beamInDotNormal, beamOutDotNormal=None, beamInDotHNormal=None):
        from . import kernels
        return kernels.crystal_amplitude_host(self, E, beamInDotNormal, beamOutDotNormal, beamInDotHNormal)


class CrystalDiamond(CrystalFcc):
    """materials/crystals_basic.py:34-80"""
    structure = 'diamond'

    def __init__(self, *args, **kwargs):
        a = kwargs.pop('a', None)
        if a is not None:
            hkl = kwargs.get('hkl', args[0] if args else (1, 1, 1))
            kwargs['d'] = a / (sum(i**2 for i in hkl))**0.5
        super().__init__(*args, **kwargs)
        self.a = a if a is not None else self.d * self.sqrthkl2

    def _diamond_to_fcc(self):
        return 1 + np.exp(0.5j * PI * sum(self.hkl))  # crystals_basic.py:77


class CrystalSi(CrystalDiamond):
    """Silicon with the thermal expansion of Swenson (materials/crystals_basic.py:83-142)."""

    def __init__(self, *args, **kwargs):
        self.a0 = 5.430710
        self.dl_l0 = self.dl_l(273.15 + 19.9)
        self.tK = kwargs.pop('tK', 297.15)
        hkl = kwargs.get('hkl', args[0] if args else (1, 1, 1))
        kwargs['hkl'] = hkl
        kwargs.pop('a', None)
        kwargs['d'] = self.get_a() / (sum(i**2 for i in hkl))**0.5
        kwargs['elements'] = 'Si'
        kwargs.setdefault('name', 'Si')
        super().__init__(*args[1:] if args else (), **kwargs)

    def dl_l(self, t=None):
        if t is None:
            t = self.tK
        if t >= 0.0 and t < 30.0:
            return -2.154537e-004
        elif t >= 30.0 and t < 130.0:
            return -2.303956e-014 * t**4 + 7.834799e-011 * t**3 - \
                1.724143e-008 * t**2 + 8.396104e-007 * t - 2.276144e-004
        elif t >= 130.0 and t < 293.0:
            return -1.223001e-011 * t**3 + 1.532991e-008 * t**2 - \
                3.263667e-006 * t - 5.217231e-005
        elif t >= 293.0 and t <= 1000.0:
            return -1.161022e-012 * t**3 + 3.311476e-009 * t**2 + \
                1.124129e-006 * t - 5.844535e-004
        else:
            return 1.0e+100

    def get_a(self):
        return self.a0 * (self.dl_l()-self.dl_l0+1)


class Multilayer(object):
    """Periodic or depth-graded multilayer on a substrate, geom 'reflected' (materials/multilayer.py:14-515): same
    constructor keywords, the same thickness laws (set_dti / set_dbi, :187-215).  The Parratt recursion over the
    2 x nPairs interfaces runs per ray in the CUDA kernels (csrc/reflect.cu: multilayer_amplitude inside
    OE.reflect, multilayer_kernel for the reference's run_parallel('get_amplitude_graded_multilayer') call)."""

    def __init__(self, tLayer=None, tThickness=0., bLayer=None, bThickness=0., nPairs=0, substrate=None,
                 tThicknessLow=0., bThicknessLow=0., idThickness=0., power=2., substRoughness=0.,
                 substThickness=np.inf, name='', geom='reflected', **kwargs):
        if bLayer is None or substrate is None or (tLayer is None and not isinstance(self, Coated)):
            raise NotImplementedError('top layer, bottom layer and substrate are all required on this path '
                                      '(a single layer on a substrate is `Coated`)')
        if geom and 'tran' in geom and not np.isfinite(substThickness):
            raise ValueError("geom='transmitted' needs a finite substThickness")
        self.tLayer, self.bLayer, self.substrate = tLayer, bLayer, substrate
        self.nPairs = int(nPairs)
        self.power = power
        self.tThicknessHigh, self.tThicknessLow = float(tThickness), float(tThicknessLow)
        self.bThicknessHigh, self.bThicknessLow = float(bThickness), float(bThicknessLow)
        self.kind = 'multilayer'
        self.geom = geom or 'reflected'
        self.idThickness = idThickness
        self.substRoughness = float(substRoughness)
        self.substThickness = substThickness
        self.name = name
        self.dti = self._graded(self.tThicknessHigh, self.tThicknessLow)
        self.dbi = self._graded(self.bThicknessHigh, self.bThicknessLow)
        self._dev = {}
        self._keep = None

    tThickness = property(lambda self: self.tThicknessHigh)
    bThickness = property(lambda self: self.bThicknessHigh)
    d = property(lambda self: float(self.tThicknessHigh + self.bThicknessHigh))

    def _graded(self, high, low):
        """set_dti / set_dbi (multilayer.py:187-215)"""
        if low:
            layers = np.arange(1, self.nPairs+1)
            qRoot = (high/low)**(1./self.power)
            qB = (self.nPairs-qRoot) / (qRoot-1.)
            qA = high * (qB+1)**self.power
            return qA * (qB+layers)**(-self.power)
        return np.ones(self.nPairs) * float(high)

    # ------------------------------------------------------------------ alignment helpers (host, multilayer.py:217-252)
    def get_sin_Bragg_angle(self, E, order=1):
        a = np.array(order * CH / (2*self.d*E))
        a[a > 1] = 1 - 1e-16
        a[a < -1] = -1 + 1e-16
        return a

    def get_Bragg_angle(self, E, order=1):
        return np.arcsin(self.get_sin_Bragg_angle(E, order))

    def get_dtheta(self, E, order=1):
        nt = self.tLayer.get_refractive_index(E).real
        nb = self.bLayer.get_refractive_index(E).real
        d_ = abs((nt-1) * self.tThicknessHigh + (nb-1) * self.bThicknessHigh) / self.d
        return self.get_Bragg_angle(E, order) - np.arcsin(((order * CH / E)**2 + self.d**2 * 8*d_)**0.5 / (2*self.d))

    # ------------------------------------------------------------------ device plumbing
    def host_tables(self):
        empty = (np.zeros(0), 0, [], [])  # no top layer (Coated)
        parts = [m.host_tables() if m is not None else empty for m in (self.tLayer, self.bLayer, self.substrate)]
        E, f1, f2, offs, lens, pos = [], [], [], [], [], 0
        for tab, n, o, ln in parts:
            E.append(tab[:n])
            f1.append(tab[n:2*n])
            f2.append(tab[2*n:])
            offs.append([pos + v for v in o])
            lens.append(list(ln))
            pos += n
        tab = np.concatenate([np.concatenate(E), np.concatenate(f1), np.concatenate(f2)])
        dti = np.ascontiguousarray(self.dti, dtype=np.float64)
        dbi = np.ascontiguousarray(self.dbi, dtype=np.float64)
        return tab, pos, (offs, dti, dbi, dti.ctypes.data, dbi.ctypes.data), lens

    def device_tables(self, device):
        import torch
        key = str(device)
        if key not in self._dev:
            tab, n, (offs, dti, dbi, _, _), lens = self.host_tables()
            ddti, ddbi = torch.from_numpy(dti).to(device), torch.from_numpy(dbi).to(device)
            self._dev[key] = (torch.from_numpy(tab).to(device), n,
                              (offs, ddti, ddbi, ddti.data_ptr(), ddbi.data_ptr()), lens)
        return self._dev[key]

    def c_struct(self, offs, lens):
        offs, dti, dbi, pdti, pdbi = offs
        ml = xrt_multilayer_params()
        for name, mat, o, ln in zip(('tLayer', 'bLayer', 'substrate'), (self.tLayer, self.bLayer, self.substrate),
                                    offs, lens):
            if mat is None:  # no top layer: nelem = 0
                setattr(ml, name, xrt_material_params())
                continue
            sub = Material.c_struct(mat, o, ln) if not hasattr(mat, 'hkl') else None
            if sub is None:
                raise NotImplementedError('crystal layers in a multilayer')
            sub.kind = 1
            setattr(ml, name, sub)
        ml.nPairs = self.nPairs
        ml.uniform = int(self.nPairs > 0 and bool(np.all(self.dti == self.dti[0]) and np.all(self.dbi == self.dbi[0])))
        ml.idThickness = float(self.idThickness)
        ml.d = self.d
        ml.dti, ml.dbi = pdti, pdbi
        ml.substRoughness = float(self.substRoughness)
        ml.transmitted = int('tran' in self.geom)
        ml.substThickness = float(self.substThickness) if np.isfinite(self.substThickness) else 0.
        ml.specular = int(self.kind == 'mirror')
        m = xrt_material_params()
        m.kind = 4
        m.multilayer = ctypes.addressof(ml)
        self._keep = (ml, dti, dbi)  # the stack and its arrays must outlive the call
        return m

    def get_amplitude(self, E, beamInDotNormal, x=None, y=None, ucl=None):
        """(ri_s, ri_p) on host arrays (multilayer.py:254-508, geom 'reflected')."""
        from . import kernels
        return kernels.material_amplitude_host(self, E, beamInDotNormal)


class Coated(Multilayer):
    """A single reflective layer on a substrate (materials/multilayer.py:528-584): a Multilayer without top layer,
    nPairs = 1, idThickness = the surface roughness, kind 'mirror' (specular reflection with the Parratt amplitude)."""

    def __init__(self, *args, **kwargs):
        coating = kwargs.pop('coating', None)
        cThickness = kwargs.pop('cThickness', 0)
        surfaceRoughness = kwargs.pop('surfaceRoughness', 0)
        Multilayer.__init__(self, *args, bLayer=coating, bThickness=cThickness, idThickness=surfaceRoughness, nPairs=1,
                            **kwargs)
        self.kind = 'mirror'

    coating = property(lambda self: self.bLayer)
    cThickness = property(lambda self: self.bThickness)
    surfaceRoughness = property(lambda self: self.idThickness)
