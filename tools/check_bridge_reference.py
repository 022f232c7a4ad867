"""Host-only check (no GPU): every element class of the UNMODIFIED reference that xrt_b200.bridge claims to adapt is
constructed from /root/reference and converted -- attribute names, material conversion, C-struct marshalling.
Run by tests/test_bridge.py when the reference tree is present (build container); prints one line per element and exits
non-zero on the first failure."""
import os
import sys
HERE = os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, "/root/reference"); sys.path.insert(0, os.path.join(HERE, "ref_stubs")); sys.path.insert(0, os.path.dirname(HERE))
import numpy as np
import xrt.backends.raycing as raycing, xrt.backends.raycing.oes as roe, xrt.backends.raycing.materials as rm
import xrt.backends.raycing.apertures as ra, xrt.backends.raycing.screens as rsc
import xrt.backends.raycing.figure_error as rfe
from xrt_b200 import bridge
raycing._VERBOSITY_ = 0
bl = raycing.BeamLine()
Au = rm.Material('Au', rho=19.3)
si = rm.CrystalSi(hkl=(1, 1, 1))
lims = dict(limPhysX=[-4, 4], limPhysY=[-100, 100])
cases = [
    roe.OE(bl, 'flat', center=(0, 1e4, 0), pitch=3e-3, material=Au, **lims),
    roe.OE(bl, 'round', center=(0, 1e4, 0), pitch=3e-3, material=Au, shape='round', **lims),
    roe.ToroidMirror(bl, 'tm', center=(0, 1e4, 0), pitch=3e-3, R=1e6, r=50., material=Au, **lims),
    roe.BentFlatMirror(bl, 'vcm', center=(0, 1e4, 0), pitch=3e-3, R=1e6, material=Au, **lims),
    roe.ConicalMirror(bl, 'cone', center=(0, 1e4, 0), pitch=3e-3, L0=1000., theta=4e-3, material=Au, **lims),
    roe.JohannCylinder(bl, 'jc', center=(0, 500, 0), pitch=0.2, Rm=2000., material=si, alpha=0.05, **lims),
    roe.JohanssonCylinder(bl, 'jsc', center=(0, 500, 0), pitch=0.2, Rm=2000., material=si, **lims),
    roe.JohannToroid(bl, 'jt', center=(0, 500, 0), pitch=0.2, Rm=2000., Rs=100., material=si, **lims),
    roe.ParabolicalMirrorParam(bl, 'pm', center=(0, 1e4, 0), pitch=3e-3, p=1e4, material=Au, **lims),
    roe.EllipticalMirrorParam(bl, 'em', center=(0, 1e4, 0), pitch=3e-3, p=1e4, q=1e3, material=Au, **lims),
    roe.LauePlate(bl, 'lp', center=(0, 1e4, 0), pitch=1.3, material=rm.CrystalSi(hkl=(1, 1, 1), geom='Laue reflected', t=0.1), **lims),
    roe.BentLaueCylinder(bl, 'blc', center=(0, 1e4, 0), pitch=1.3, R=2e4, material=rm.CrystalSi(hkl=(1, 1, 1), geom='Laue reflected', t=0.1), **lims),
    roe.NormalFZP(bl, 'fzp', center=(0, 1e4, 0), pitch=np.pi/2, material=rm.Material('Au', rho=19.3, kind='FZP'), f=100., E=9000., N=100),
    roe.BlazedGrating(bl, 'bg', center=(0, 1e4, 0), pitch=0.03, material=rm.Material('Au', rho=19.3, kind='mirror'), blaze=0.02, rho=300., **lims),
    roe.OE(bl, 'coated', center=(0, 1e4, 0), pitch=3e-3, material=rm.Coated(coating=rm.Material('Pt', rho=21.45), cThickness=300., surfaceRoughness=3., substrate=rm.Material('Si', rho=2.33), substRoughness=4.), **lims),
    roe.OE(bl, 'ml', center=(0, 1e4, 0), pitch=0.02, material=rm.Multilayer(rm.Material('W', rho=19.3), 12., rm.Material('Si', rho=2.33), 22., 60, rm.Material('Si', rho=2.33), idThickness=2.), **lims),
    roe.OE(bl, 'fe', center=(0, 1e4, 0), pitch=3e-3, material=Au, figureError=rfe.GaussianBump(bumpHeight=10., limPhysX=[-4, 4], limPhysY=[-100, 100], gridStep=4.), **lims),
    roe.DCM(bl, 'dcm', center=(0, 2e4, 0), bragg=0.22, material=si, material2=si, cryst2perpTransl=10., limPhysX=[-50, 50], limPhysY=[-100, 100], limPhysX2=[-50, 50], limPhysY2=[-100, 200]),
]
for oe in cases:
    try:
        out = bridge.convert_oe(oe)
        p = out._make_params(out.pitch, out.roll + out.positionRoll, out.yaw, out.dx) if not hasattr(out, 'bragg') and out.figureError is None and not hasattr(out, 'rn') else None
        print('%-8s -> %-24s kind %s surf %s' % (oe.name, type(out).__name__, getattr(out.material, 'kind', None), None if p is None else p.surface))
    except Exception as e:
        print('%-8s FAILED: %s: %s' % (oe.name, type(e).__name__, e))
        sys.exit(1)
