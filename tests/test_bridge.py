"""Method-level drop-in: xrt-shaped host objects (duck-typed with xrt's attribute names, filled from the golden
files the reference produced) go through xrt_b200.bridge and come back as host beams equal to the reference's."""
import types

import numpy as np
import pytest

from conftest import golden
from helpers import oracle_beam, compare_beams, BEAM_FIELDS


class ToroidMirror(object):  # class name is what the adapter dispatches on, like isinstance in xrt
    pass


def _fake_xrt_oe(g):
    class Elem(object):
        name = 'Au'
    mat = types.SimpleNamespace(elements=[Elem()], quantities=[1.], kind='mirror', rho=float(g['rho']), t=None,
                                table='Chantler total')
    oe = ToroidMirror()
    bl = types.SimpleNamespace(sinAzimuth=float(g['sinAzimuth']), cosAzimuth=float(g['cosAzimuth']), oes=[])
    oe.__dict__.update(dict(
        bl=bl, name='TM', center=list(g['center']), pitch=float(g['pitch']), roll=float(g['roll']),
        positionRoll=float(g['positionRoll']), yaw=float(g['yaw']), rotationSequence=str(g['rotationSequence']),
        extraPitch=float(g['extraPitch']), extraRoll=float(g['extraRoll']), extraYaw=float(g['extraYaw']),
        extraRotationSequence='RzRyRx', alpha=None, surfPhysX=list(g['surfPhysX']), surfPhysY=list(g['surfPhysY']),
        surfOptX=None, surfOptY=None, overEdge=str(g['overEdge']), shape='rect', material=mat,
        lostNum=int(g['lostNum']), dx=0, R=float(g['R']), r=float(g['r']), footprint=[]))
    return oe


@pytest.mark.gpu
def test_bridge_reflect_host_objects(cuda_device):
    from xrt_b200 import bridge
    g = golden('reflect_toroid_Au')
    oe = _fake_xrt_oe(g)
    gb, lb = bridge.reflect(oe, oracle_beam(g, 'in_'))
    assert isinstance(gb.x, np.ndarray)
    bad = compare_beams(gb.__dict__, {f: g['gb_' + f] for f in BEAM_FIELDS}, BEAM_FIELDS, 1e-12, 'gb_')
    bad += compare_beams(lb.__dict__, {f: g['lb_' + f] for f in BEAM_FIELDS + ('theta',)},
                         BEAM_FIELDS + ('theta',), 1e-12, 'lb_')
    assert not bad, bad
    np.testing.assert_allclose(oe.footprint[0], g['footprint'], rtol=1e-12, atol=1e-12)
    assert oe.bl.oes == []  # the adapter leaves the user's beamline numbering alone


class Plate(object):
    pass


class DoubleParaboloidLens(Plate):
    pass


class OE(object):
    pass


def _fake_common(g, cls, mat, **extra):
    oe = cls()
    bl = types.SimpleNamespace(sinAzimuth=float(g['sinAzimuth']), cosAzimuth=float(g['cosAzimuth']), oes=[])
    oe.__dict__.update(dict(
        bl=bl, name='E', center=list(g['center']), pitch=float(g['pitch']), roll=float(g['roll']),
        positionRoll=float(g['positionRoll']), yaw=float(g['yaw']), rotationSequence=str(g['rotationSequence']),
        extraPitch=float(g['extraPitch']), extraRoll=float(g['extraRoll']), extraYaw=float(g['extraYaw']),
        extraRotationSequence='RzRyRx', alpha=None, surfPhysX=list(g['surfPhysX']), surfPhysY=list(g['surfPhysY']),
        surfOptX=None, surfOptY=None, overEdge=str(g['overEdge']), shape='rect', material=mat,
        lostNum=int(g['lostNum']), dx=0, footprint=[]))
    oe.__dict__.update(extra)
    return oe


def _fake_material(name, g, kind, **kw):
    class Elem(object):
        pass
    e = Elem()
    e.name = name
    return types.SimpleNamespace(elements=[e], quantities=[1.], kind=kind, rho=float(g['rho']), t=None,
                                 table='Chantler total', **kw)


@pytest.mark.gpu
def test_bridge_refract_and_grating_host_objects(cuda_device):
    from xrt_b200 import bridge
    g = golden('refract_crl_double')
    lens = _fake_common(g, DoubleParaboloidLens, _fake_material('Be', g, 'lens'), t=float(g['t']), wedgeAngle=0.,
                        focus=float(g['focus']), zmax=float(g['zmax']), nCRL=int(g['nCRL']))
    gb, lo1, lo2 = bridge.multiple_refract(lens, oracle_beam(g, 'in_'))
    bad = []
    for pre, b in (('gb_', gb), ('lo1_', lo1), ('lo2_', lo2)):
        bad += compare_beams(b.__dict__, {f: g[pre + f] for f in BEAM_FIELDS}, BEAM_FIELDS, 1e-12, pre)
    assert not bad, bad
    g = golden('reflect_grating_vls')
    gd = [str(g['gratingAxis'])] + [float(v) for v in g['gratingDensity']]
    oe = _fake_common(g, OE, _fake_material('Au', g, 'grating', efficiency=None), gratingDensity=gd,
                      order=int(g['order']))
    gb, lb = bridge.reflect(oe, oracle_beam(g, 'in_'))
    bad = compare_beams(gb.__dict__, {f: g['gb_' + f] for f in BEAM_FIELDS}, BEAM_FIELDS, 1e-12, 'gb_')
    bad += compare_beams(lb.__dict__, {f: g['lb_' + f] for f in BEAM_FIELDS}, BEAM_FIELDS, 1e-12, 'lb_')
    assert not bad, bad


class BentFlatMirror(OE):
    pass


class VCM(BentFlatMirror):  # a subclass with a surface of its own must not be silently taken for its base
    pass


@pytest.mark.gpu
def test_bridge_spliced_element_and_multiple_reflect(cuda_device):
    """An xrt-shaped BentFlatMirror goes through the spliced-surface plug-in: multiple_reflect (8 whispering-gallery
    passes) equals the reference's beams; an unknown subclass is refused instead of being traced as something else."""
    from xrt_b200 import bridge
    g = golden('multi_reflect_bentflat')
    oe = _fake_common(g, BentFlatMirror, _fake_material('Au', g, 'mirror'), R=float(g['R']),
                      limPhysY=list(g['surfPhysY']))
    gb, lbN = bridge.multiple_reflect(oe, oracle_beam(g, 'in_'), maxReflections=8)
    bad = compare_beams(gb.__dict__, {f: g['gb_' + f] for f in BEAM_FIELDS}, BEAM_FIELDS, 1e-12, 'gb_')
    bad += compare_beams(lbN.__dict__, {f: g['lbN_' + f] for f in BEAM_FIELDS}, BEAM_FIELDS, 1e-12, 'lbN_')
    assert not bad, bad
    assert np.array_equal(gb.nRefl, g['gb_nRefl']) and len(oe.footprint) == 8
    other = _fake_common(g, VCM, _fake_material('Au', g, 'mirror'), R=float(g['R']), limPhysY=list(g['surfPhysY']))
    with pytest.raises(NotImplementedError):
        bridge.reflect(other, oracle_beam(g, 'in_'))


@pytest.mark.gpu
def test_bridge_figure_error_host_objects(cuda_device):
    """An xrt-shaped toroid carrying an xrt-shaped FigureError (a scipy spline + shifts) is traced by the figure-error
    plug-in variant: states equal the reference's, positions to 1e-12."""
    from scipy import interpolate
    from xrt_b200 import bridge
    g = golden('reflect_figerr_toroid_bump')
    oe = _fake_xrt_oe(g)
    oe.figureError = types.SimpleNamespace(
        local_z_spline=interpolate.RectBivariateSpline(g['fe_y1d'], g['fe_x1d'], g['fe_z'], kx=3, ky=3),
        xShift=float(g['fe_xShift']), yShift=float(g['fe_yShift']), name='bump')
    gb, lb = bridge.reflect(oe, oracle_beam(g, 'in_'))
    assert np.array_equal(gb.state, g['gb_state'])
    geo = ('x', 'y', 'z', 'a', 'b', 'c', 'Jss', 'Jpp', 'Jsp')
    bad = compare_beams(gb.__dict__, {f: g['gb_' + f] for f in BEAM_FIELDS}, geo, 1e-12, 'gb_')
    assert not bad, bad


@pytest.mark.gpu
def test_bridge_propagate_round_aperture_host_objects(cuda_device):
    from xrt_b200 import bridge
    g = golden('propagate_round')
    ap = types.SimpleNamespace(name='pipe', center=list(g['center']), r=float(g['r']), shape='round',
                               isBeamStop=False, xyz=[list(g['ex']), list(g['ey']), list(g['ez'])],
                               lostNum=int(g['lostNum']))
    beam = oracle_beam(g, 'in_')
    lo = bridge.propagate(ap, beam)
    assert not compare_beams(lo.__dict__, {f: g['out_' + f] for f in BEAM_FIELDS}, BEAM_FIELDS, 1e-12)
    assert np.array_equal(beam.state, g['in_state_after'])


@pytest.mark.skipif(not __import__('os').path.isdir('/root/reference'), reason='needs the reference tree')
def test_bridge_converts_real_reference_objects():
    """Build container only (no GPU): the unmodified reference's own element / material / figure-error objects go
    through bridge.convert_oe -- 18 classes -- in a separate interpreter (tools/check_bridge_reference.py)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, 'tools', 'check_bridge_reference.py')],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True, timeout=600)
    assert res.returncode == 0, res.stdout
    lines = [ln for ln in res.stdout.splitlines() if ' -> ' in ln]
    assert len(lines) == 18 and not any('FAILED' in ln for ln in res.stdout.splitlines()), res.stdout
