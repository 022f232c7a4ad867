"""Shared builders: oracle OESpec / product OE from the scalars stored in a golden file, beam conversion."""
import numpy as np

from oracle import reflect as orf, beam as ob, materials as om

BEAM_FIELDS = ('x', 'y', 'z', 'a', 'b', 'c', 'path', 'E', 'Jss', 'Jpp', 'Jsp', 'state', 'Es', 'Ep')


def oracle_spec(g, surface='flat', material=None, **kw):
    return orf.OESpec(
        azimuth=(float(g['sinAzimuth']), float(g['cosAzimuth'])), center=list(g['center']),
        pitch=float(g['pitch']), roll=float(g['roll']) + float(g['positionRoll']), yaw=float(g['yaw']),
        rotationSequence=str(g['rotationSequence']), extraPitch=float(g['extraPitch']),
        extraRoll=float(g['extraRoll']), extraYaw=float(g['extraYaw']), dx=float(g['dx']),
        surfPhysX=list(g['surfPhysX']), surfPhysY=list(g['surfPhysY']),
        surfOptX=list(g['surfOptX']) if 'surfOptX' in g.files else None,
        surfOptY=list(g['surfOptY']) if 'surfOptY' in g.files else None,
        overEdge=str(g['overEdge']), lostNum=int(g['lostNum']), surface=surface, material=material, **kw)


def product_beamline(g):
    from xrt_b200.beamline import BeamLine
    bl = BeamLine()
    bl.sinAzimuth = float(g['sinAzimuth'])
    bl.cosAzimuth = float(g['cosAzimuth'])
    return bl


def product_oe_kwargs(g):
    kw = dict(center=list(g['center']), pitch=float(g['pitch']), roll=float(g['roll']),
              positionRoll=float(g['positionRoll']), yaw=float(g['yaw']),
              rotationSequence=str(g['rotationSequence']), extraPitch=float(g['extraPitch']),
              extraRoll=float(g['extraRoll']), extraYaw=float(g['extraYaw']),
              limPhysX=list(g['surfPhysX']), limPhysY=list(g['surfPhysY']), overEdge=str(g['overEdge']))
    if 'surfOptX' in g.files:
        kw['limOptX'] = list(g['surfOptX'])
    if 'surfOptY' in g.files:
        kw['limOptY'] = list(g['surfOptY'])
    return kw


def oracle_beam(g, prefix):
    return ob.Beam.from_npz(g, prefix)


def device_beam(b):
    """oracle Beam / dict of arrays -> xrt_b200 device Beam."""
    from xrt_b200.beams import Beam
    d = b if isinstance(b, dict) else b.__dict__
    return Beam.from_numpy({k: v for k, v in d.items() if isinstance(v, np.ndarray)})


def max_rel(a, b, scale=None):
    a = np.asarray(a)
    b = np.asarray(b)
    if scale is None:
        scale = max(np.abs(b).max(), 1e-300)
    return np.abs(a - b).max() / scale


GROUPS = (('x', 'y', 'z'), ('a', 'b', 'c'), ('Jss', 'Jpp', 'Jsp'), ('Es', 'Ep'), ('s', 'r'))


def compare_beams(dev, ref, fields, rtol, what=''):
    """dev: dict from Beam.to_numpy(); ref: object/dict with numpy arrays.  Every field is compared with an
    absolute tolerance of rtol x the natural scale of its group (positions: max |x,y,z|; directions: 1;
    coherency matrix: max J; amplitudes: max |E|), so that components that are zero up to rounding (e.g. Jpp
    of a horizontally polarised beam after a frame rotation and back) are judged against the beam, not
    against themselves.  Returns the list of failures."""
    bad = []
    rd = ref if isinstance(ref, dict) else ref.__dict__
    scales = {}
    for grp in GROUPS:
        sc = 0.
        for f in grp:
            if f in rd and isinstance(rd[f], np.ndarray) and rd[f].size:
                sc = max(sc, float(np.abs(rd[f]).max()))
        for f in grp:
            scales[f] = max(sc, 1e-300)
    for f in fields:
        if f not in rd or f not in dev:
            continue
        if f == 'state':
            if not np.array_equal(dev[f], rd[f]):
                bad.append((what + f, int((dev[f] != rd[f]).sum())))
            continue
        if rd[f].size == 0:
            continue
        scale = scales.get(f, max(float(np.abs(rd[f]).max()), 1e-300))
        err = np.abs(dev[f] - rd[f]).max() / scale
        if not err <= rtol:
            bad.append((what + f, float(err)))
    return bad
