# -*- coding: utf-8 -*-
"""Adapters that let *xrt's own objects* (duck-typed: anything with xrt's attribute names) run on the CUDA
kernels: the method-level binding sketched in INTEGRATION.md section 2.

    bridge.reflect(xrt_oe, xrt_beam, needLocal=True)          -> (gb, lb)      OE.reflect, oes/reflect.py:18
    bridge.multiple_reflect(xrt_oe, xrt_beam, maxReflections) -> (gb, lbN)     OE.multiple_reflect, oes/reflect.py:165
    bridge.double_reflect(xrt_dcm, xrt_beam, needLocal=True)  -> (gb, lo1, lo2)  DCM.double_reflect, dcm.py:248
    bridge.expose(xrt_screen, xrt_beam)                       -> blo           Screen / HemisphericScreen.expose
    bridge.propagate(xrt_aperture, xrt_beam)                  -> lo            Rectangular / RoundAperture.propagate
    bridge.double_refract(xrt_plate, xrt_beam)                -> (gb, lo1, lo2)  Plate.double_refract, refractive.py:171
    bridge.multiple_refract(xrt_lens, xrt_beam)               -> (gb, lo1, lo2)  multiple_refract, refractive.py:457

Inputs and outputs are host objects with numpy attribute arrays (what xrt's flow, `get_output()` and plots
consume), so every call pays one H2D and one D2H of the Beam; pipelines that stay on the GPU use the classes of
xrt_b200.oes / planes / sources directly."""
import numpy as np

from . import materials as _mat, oes as _oes, planes as _planes
from .beams import Beam


class HostBeam(object):
    """Attribute bag with numpy arrays, like xrt's Beam."""

    def __init__(self, d):
        self.__dict__.update(d)


def _device_beam(b):
    d = {k: v for k, v in b.__dict__.items() if isinstance(v, np.ndarray) and v.ndim == 1 and
         k in ('x', 'y', 'z', 'a', 'b', 'c', 'path', 'E', 'Jss', 'Jpp', 'Jsp', 'state', 'Es', 'Ep')}
    out = Beam.from_numpy(d)
    if getattr(b, 'createdByDiffract', False):
        out.createdByDiffract = True
    return out


def convert_material(m):
    """xrt Material / CrystalFcc / CrystalDiamond / CrystalSi instance -> xrt_b200 material."""
    if m is None:
        return None
    if hasattr(m, 'bLayer'):  # Multilayer / Coated (materials/multilayer.py:14-584)
        if 'tran' in (getattr(m, 'geom', None) or 'reflected'):
            raise NotImplementedError("geom='transmitted' multilayers are outside the hot path")
        common = dict(substrate=convert_material(m.substrate), substRoughness=m.substRoughness)
        if m.tLayer is None:
            return _mat.Coated(coating=convert_material(m.bLayer), cThickness=m.bThickness,
                               surfaceRoughness=m.idThickness, **common)
        return _mat.Multilayer(tLayer=convert_material(m.tLayer), tThickness=m.tThickness,
                               bLayer=convert_material(m.bLayer), bThickness=m.bThickness, nPairs=m.nPairs,
                               tThicknessLow=m.tThicknessLow, bThicknessLow=m.bThicknessLow,
                               idThickness=m.idThickness, power=m.power, **common)
    if not hasattr(m, 'elements'):  # EmptyMaterial (materials/__init__.py:101-113)
        return _mat.EmptyMaterial(kind=m.kind)
    names = [e.name for e in m.elements]
    table = getattr(m, 'table', 'Chantler total')
    if getattr(m, 'kind', None) == 'crystal':
        cls = type(m).__name__
        kw = dict(hkl=tuple(m.hkl), d=m.d, V=getattr(m, 'V', None), elements=names[0], factDW=m.factDW,
                  geom=m.geom, table=table, t=m.t)
        if 'Fcc' in cls:
            return _mat.CrystalFcc(**kw)
        if any('Diamond' in c.__name__ for c in type(m).__mro__):
            return _mat.CrystalDiamond(**kw)
        raise NotImplementedError('crystal class {0} is outside the hot path'.format(cls))
    return _mat.Material(names, list(m.quantities), kind=m.kind, rho=m.rho, t=m.t, table=table,
                         efficiency=getattr(m, 'efficiency', None))


def _oe_kwargs(oe):
    kw = dict(center=list(oe.center), pitch=float(oe.pitch), roll=float(oe.roll),
              positionRoll=float(oe.positionRoll), yaw=float(oe.yaw), rotationSequence=oe.rotationSequence,
              extraPitch=float(oe.extraPitch), extraRoll=float(oe.extraRoll), extraYaw=float(oe.extraYaw),
              extraRotationSequence=oe.extraRotationSequence, alpha=getattr(oe, 'alpha', None),
              limPhysX=list(oe.surfPhysX), limPhysY=list(oe.surfPhysY),
              limOptX=None if oe.surfOptX is None else list(oe.surfOptX),
              limOptY=None if oe.surfOptY is None else list(oe.surfOptY),
              overEdge=getattr(oe, 'overEdge', 'yMax'), shape=oe.shape, material=convert_material(oe.material))
    fe = getattr(oe, 'figureError', None)
    if fe is not None and hasattr(fe, 'local_z_spline'):  # any of xrt's FigureError classes (figure_error.py:207-212)
        from .figure_error import FigureErrorSpline
        kw['figureError'] = FigureErrorSpline(fe.local_z_spline, fe.xShift, fe.yShift, getattr(fe, 'name', ''))
    if getattr(oe, 'gratingDensity', None) is not None:
        kw['gratingDensity'] = list(oe.gratingDensity)
    if (getattr(kw['material'], 'kind', None) in ('grating', 'auto') or 'gratingDensity' in kw) and \
            type(oe).__name__ != 'NormalFZP':
        kw['order'] = getattr(oe, 'order', 1)
    return kw


def convert_oe(oe):
    """xrt OE / ToroidMirror / EllipticalMirrorParam / DCM instance -> the xrt_b200 element."""
    cls = [c.__name__ for c in type(oe).__mro__]
    kw = _oe_kwargs(oe)
    if 'Plate' in cls:  # Plate and the parabolic lenses (oes/refractive.py)
        kw.update(t=float(oe.t), wedgeAngle=float(getattr(oe, 'wedgeAngle', 0)))
        for name in ('DoubleParabolicCylinderLens', 'DoubleParaboloidLens', 'ParabolicCylinderFlatLens',
                     'ParaboloidFlatLens'):
            if name in cls:
                out = getattr(_oes, name)(oe.bl, oe.name, focus=float(oe.focus), zmax=oe.zmax, nCRL=int(oe.nCRL),
                                          **kw)
                break
        else:
            out = _oes.Plate(oe.bl, oe.name, **kw)
    elif 'DCM' in cls:
        for k in ('bragg', 'cryst1roll', 'cryst2roll', 'cryst2pitch', 'cryst2finePitch', 'cryst2perpTransl',
                  'cryst2longTransl'):
            kw[k] = float(getattr(oe, k))
        kw.update(limPhysX2=list(oe.surfPhysX2), limPhysY2=list(oe.surfPhysY2),
                  limOptX2=None if oe.surfOptX2 is None else list(oe.surfOptX2),
                  limOptY2=None if oe.surfOptY2 is None else list(oe.surfOptY2),
                  material2=convert_material(oe.material2))
        out = _oes.DCM(oe.bl, oe.name, **kw)
    elif 'ToroidMirror' in cls:
        out = _oes.ToroidMirror(oe.bl, oe.name, R=float(oe.R), r=float(oe.r), **kw)
    elif 'EllipticalMirrorParam' in cls:
        out = _oes.EllipticalMirrorParam(oe.bl, oe.name, p=oe.p, q=oe.q, isCylindrical=oe.isCylindrical, **kw)
        for k in ('ellipseA', 'ellipseB', 'y0', 'z0', 'cosGamma', 'sinGamma'):
            setattr(out, k, float(getattr(oe, k)))  # take xrt's own ellipse parameters verbatim
    elif 'ParabolicalMirrorParam' in cls:
        if getattr(oe, 'isClosed', False):
            raise NotImplementedError('isClosed=True is outside the hot path')
        pq = dict(p=oe.p) if oe.p is not None else dict(q=oe.q)
        out = _oes.ParabolicalMirrorParam(oe.bl, oe.name, isCylindrical=oe.isCylindrical, **pq, **kw)
        for k in ('parabParam', 'y0', 'z0', 'cosGamma', 'sinGamma'):
            setattr(out, k, float(getattr(oe, k)))  # xrt's own parabola parameters verbatim
    elif 'JohanssonCylinder' in cls or 'JohannCylinder' in cls:
        out = getattr(_oes, 'JohanssonCylinder' if 'JohanssonCylinder' in cls else 'JohannCylinder')(
            oe.bl, oe.name, Rm=float(oe.Rm), crossSection=oe.crossSection, **kw)
    elif 'JohannToroid' in cls and not any(c in cls for c in ('JohanssonToroid', 'GeneralBraggToroid', 'DicedOE')):
        out = _oes.JohannToroid(oe.bl, oe.name, Rm=float(oe.Rm), Rs=float(oe.Rs), **kw)
    elif type(oe).__name__ == 'BlazedGrating':
        if getattr(oe, 'gratingDensity', None) is not None:
            raise NotImplementedError('a variable line density is outside the hot path')
        kw.pop('gratingDensity', None), kw.pop('order', None)
        out = _oes.BlazedGrating(oe.bl, oe.name, blaze=float(oe.blaze), antiblaze=float(oe.antiblaze),
                                 rho=float(oe.rho0), **kw)
    elif type(oe).__name__ == 'NormalFZP':
        kw.pop('limPhysX'), kw.pop('limPhysY')
        out = _oes.NormalFZP(oe.bl, oe.name, f=oe.f, E=oe.E, N=oe.N, isCentralZoneBlack=oe.isCentralZoneBlack,
                             order=getattr(oe, 'order', 1), **kw)
    elif type(oe).__name__ == 'LauePlate':
        out = _oes.LauePlate(oe.bl, oe.name, **kw)
    elif type(oe).__name__ == 'BentLaueCylinder':
        if getattr(oe.material, 'volumetricDiffraction', False):
            raise NotImplementedError('volumetric diffraction is outside the hot path')
        out = _oes.BentLaueCylinder(oe.bl, oe.name, R=float(oe.R), crossSection=oe.crossSection, **kw)
    elif 'ConicalMirror' in cls:
        out = _oes.ConicalMirror(oe.bl, oe.name, L0=float(oe.L0), theta=float(oe.theta), **kw)
    elif 'BentFlatMirror' in cls and type(oe).__name__ in ('BentFlatMirror', 'SimpleVCM'):
        out = _oes.BentFlatMirror(oe.bl, oe.name, R=float(oe.R), **kw)
        out.limPhysY = list(oe.limPhysY)  # local_z reads limPhysY[0], not surfPhysY (oes/__init__.py:293)
    elif type(oe).__name__ == 'OE':
        out = _oes.OE(oe.bl, oe.name, **kw)
    else:
        raise NotImplementedError('{0} is outside the hot path'.format(type(oe).__name__))
    if oe.bl is not None and out in getattr(oe.bl, 'oes', []):
        oe.bl.oes.remove(out)  # the adapter must not renumber the user's beamline
    out.lostNum = oe.lostNum
    out.dx = getattr(oe, 'dx', 0)
    return out


def reflect(oe, beam, needLocal=True):
    dev = convert_oe(oe)
    gb, lb = dev.reflect(_device_beam(beam), needLocal=needLocal, withFootprint=True)
    oe.footprint = dev.footprint
    hgb = HostBeam(gb.to_numpy())
    return hgb, (HostBeam(lb.to_numpy()) if needLocal else hgb)


def multiple_reflect(oe, beam, maxReflections=1000):
    dev = convert_oe(oe)
    gb, lbN = dev.multiple_reflect(_device_beam(beam), maxReflections=maxReflections)
    oe.footprint = dev.footprint
    return HostBeam(gb.to_numpy()), HostBeam(lbN.to_numpy())


def double_reflect(dcm, beam, needLocal=True):
    dev = convert_oe(dcm)
    gb, lo1, lo2 = dev.double_reflect(_device_beam(beam), needLocal=needLocal)
    hgb = HostBeam(gb.to_numpy())
    if not needLocal:
        return hgb, hgb, hgb
    return hgb, HostBeam(lo1.to_numpy()), HostBeam(lo2.to_numpy())


def propagate(aperture, beam):
    """RectangularAperture.propagate / RoundAperture.propagate (apertures.py:333-413, :769-825): returns the local
    beam; `beam.state` of the caller's host beam is marked like the reference marks it."""
    if getattr(aperture, 'shape', '') == 'round' or hasattr(aperture, 'r'):
        ap = _planes.RoundAperture(None, aperture.name, center=list(aperture.center), r=float(aperture.r),
                                   isBeamStop=getattr(aperture, 'isBeamStop', False))
    else:
        ap = _planes.RectangularAperture(None, aperture.name, center=list(aperture.center),
                                         blades=dict(aperture.blades),
                                         isBeamStop=getattr(aperture, 'isBeamStop', False))
    ap.x, ap.y, ap.z = [list(v) for v in aperture.xyz]
    ap.lostNum = aperture.lostNum
    dev = _device_beam(beam)
    lo = ap.propagate(dev)
    beam.state[:] = dev.state.cpu().numpy()
    return HostBeam(lo.to_numpy())


def expose(screen, beam, onlyPositivePath=False):
    if hasattr(screen, 'thetaOffset'):  # HemisphericScreen (screens.py:422-545)
        scr = _planes.HemisphericScreen(None, screen.name, center=list(screen.center), R=float(screen.R),
                                        x=list(screen.x), z=list(screen.z), phiOffset=screen.phiOffset,
                                        thetaOffset=screen.thetaOffset)
        scr.y = list(screen.y)
        scr.lostNum = screen.lostNum
        return HostBeam(scr.expose(_device_beam(beam), onlyPositivePath=onlyPositivePath).to_numpy())
    scr = _planes.Screen(None, screen.name, center=list(screen.center), compressX=screen.compressX,
                         compressZ=screen.compressZ)
    scr.x, scr.y, scr.z = list(screen.x), list(screen.y), list(screen.z)
    scr.lostNum = screen.lostNum
    blo = scr.expose(_device_beam(beam), onlyPositivePath=onlyPositivePath, withFootprint=True)
    screen.footprint = scr.footprint
    return HostBeam(blo.to_numpy())


def _three(gb, lo1, lo2, needLocal):
    hgb = HostBeam(gb.to_numpy())
    if not needLocal:
        return hgb, hgb, hgb
    return hgb, HostBeam(lo1.to_numpy()), HostBeam(lo2.to_numpy())


def double_refract(plate, beam, needLocal=True):
    dev = convert_oe(plate)
    return _three(*dev.double_refract(_device_beam(beam), needLocal=needLocal), needLocal)


def multiple_refract(lens, beam, needLocal=True):
    dev = convert_oe(lens)
    return _three(*dev.multiple_refract(_device_beam(beam), needLocal=needLocal), needLocal)
