"""Oracle beam container and frame transforms (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Restates xrt/backends/raycing/sources/beams.py:96-186 (Beam), :409-445 (copy_beam), :448-479
(rotate_coherency_matrix), _rotate.py:5-57 and beamline.py:230-287 (global <-> virgin local)."""
import numpy as np

FIELDS = ('x', 'y', 'z', 'a', 'b', 'c', 'path', 'E', 'Jss', 'Jpp', 'Jsp', 'state', 'Es', 'Ep')


COPIED = ('x', 'y', 'z', 'sourceSIGMAx', 'sourceSIGMAz', 'filamentDX', 'filamentDZ', 'filamentDtheta', 'filamentDpsi',
          'filamentDgamma', 'state', 'a', 'b', 'c', 'path', 'E', 'Jss', 'Jpp', 'Jsp', 'elevationD', 'elevationX',
          'elevationY', 'elevationZ', 's', 'phi', 'r', 'theta', 'order', 'accepted', 'acceptedE', 'seeded', 'seededI',
          'Es', 'Ep', 'nRefl', 'sourceWeight')


class Beam(object):
    def __init__(self, nrays=0, copyFrom=None, withAmplitudes=False):
        if copyFrom is not None:  # only the listed attributes travel (beams.py:96-117): no area, toOE, ...
            for k in COPIED:
                if hasattr(copyFrom, k):
                    setattr(self, k, np.copy(getattr(copyFrom, k)))
            return
        self.x = np.zeros(nrays)
        self.y = np.zeros(nrays)
        self.z = np.zeros(nrays)
        self.state = np.zeros(nrays, dtype=np.int32)
        self.a = np.zeros(nrays)
        self.b = np.ones(nrays)
        self.c = np.zeros(nrays)
        self.path = np.zeros(nrays)
        self.E = np.ones(nrays) * 9000.
        self.Jss = np.ones(nrays)
        self.Jpp = np.zeros(nrays)
        self.Jsp = np.zeros(nrays, dtype=complex)
        if withAmplitudes:
            self.Es = np.zeros(nrays, dtype=complex)
            self.Ep = np.zeros(nrays, dtype=complex)

    @classmethod
    def from_npz(cls, g, prefix):
        b = cls.__new__(cls)
        for key in g.files if hasattr(g, 'files') else g.keys():
            if key.startswith(prefix):
                setattr(b, key[len(prefix):], np.array(g[key]))
        return b

    def filter_by_index(self, indarr):
        for k, v in list(self.__dict__.items()):
            if isinstance(v, np.ndarray):
                setattr(self, k, v[indarr])


def concatenate(self, beam):
    """Beam.concatenate (beams.py:230-295): the geometric / coherency fields and the optional per-ray extras
    (nRefl, s, phi, r, theta, order, Es, Ep) are stacked."""
    for k in ('state', 'x', 'y', 'z', 'a', 'b', 'c', 'path', 'E', 'Jss', 'Jpp', 'Jsp'):
        setattr(self, k, np.concatenate((getattr(self, k), getattr(beam, k))))
    for k in ('nRefl', 'elevationD', 'elevationX', 'elevationY', 'elevationZ', 's', 'phi', 'r', 'theta', 'order', 'Es', 'Ep'):
        if hasattr(self, k) and hasattr(beam, k):
            setattr(self, k, np.concatenate((getattr(self, k), getattr(beam, k))))


def rotate_x(y, z, cosangle, sinangle):
    return cosangle*y - sinangle*z, sinangle*y + cosangle*z


def rotate_y(x, z, cosangle, sinangle):
    return cosangle*x + sinangle*z, -sinangle*x + cosangle*z


def rotate_z(x, y, cosangle, sinangle):
    return cosangle*x - sinangle*y, sinangle*x + cosangle*y


def rotate_beam(beam, indarr=None, rotationSequence='RzRyRx', pitch=0, roll=0, yaw=0):
    angles = {'z': yaw, 'y': roll, 'x': pitch}
    rotates = {'z': rotate_z, 'y': rotate_y, 'x': rotate_x}
    coords1 = {'z': beam.x, 'y': beam.x, 'x': beam.y}
    coords2 = {'z': beam.y, 'y': beam.z, 'x': beam.z}
    vcomps1 = {'z': beam.a, 'y': beam.a, 'x': beam.b}
    vcomps2 = {'z': beam.b, 'y': beam.c, 'x': beam.c}
    if rotationSequence[0] == '-':
        seq = rotationSequence[6] + rotationSequence[4] + rotationSequence[2]
    else:
        seq = rotationSequence[1] + rotationSequence[3] + rotationSequence[5]
    for s in seq:
        angle, rotate = angles[s], rotates[s]
        c1, c2 = coords1[s], coords2[s]
        v1, v2 = vcomps1[s], vcomps2[s]
        if angle != 0:
            cA = np.cos(angle)
            sA = np.sin(angle)
            if indarr is None:
                indarr = slice(None)
            c1[indarr], c2[indarr] = rotate(c1[indarr], c2[indarr], cA, sA)
            v1[indarr], v2[indarr] = rotate(v1[indarr], v2[indarr], cA, sA)


def global_to_virgin_local(azimuth_sc, beam, lo, center=None, part=None):
    """azimuth_sc = (sinAzimuth, cosAzimuth) of the BeamLine, or a 3x3 basis (lx, ly, lz) as for screens."""
    if part is None:
        part = np.ones(beam.x.shape, dtype=bool)
    if center is None:
        center = [0, 0, 0]
    lo.x[part] = beam.x[part] - center[0]
    lo.y[part] = beam.y[part] - center[1]
    lo.z[part] = beam.z[part] - center[2]
    if len(azimuth_sc) == 2:
        a0, b0 = azimuth_sc
        if a0 == 0:
            lo.a[part] = beam.a[part]
            lo.b[part] = beam.b[part]
        else:
            lo.x[part], lo.y[part] = rotate_z(lo.x[part], lo.y[part], b0, a0)
            lo.a[part], lo.b[part] = rotate_z(beam.a[part], beam.b[part], b0, a0)
        lo.c[part] = beam.c[part]
    else:
        lx, ly, lz = azimuth_sc
        xyz = lo.x[part], lo.y[part], lo.z[part]
        lo.x[part], lo.y[part], lo.z[part] = (
            sum(c*b for c, b in zip(lx, xyz)),
            sum(c*b for c, b in zip(ly, xyz)),
            sum(c*b for c, b in zip(lz, xyz)))
        abc = beam.a[part], beam.b[part], beam.c[part]
        lo.a[part], lo.b[part], lo.c[part] = (
            sum(c*b for c, b in zip(lx, abc)),
            sum(c*b for c, b in zip(ly, abc)),
            sum(c*b for c, b in zip(lz, abc)))


def virgin_local_to_global(azimuth_sc, vlb, center=None, part=None):
    if part is None:
        part = np.ones(vlb.x.shape, dtype=bool)
    a0, b0 = azimuth_sc
    if a0 != 0:
        vlb.a[part], vlb.b[part] = rotate_z(vlb.a[part], vlb.b[part], b0, -a0)
        vlb.x[part], vlb.y[part] = rotate_z(vlb.x[part], vlb.y[part], b0, -a0)
    if center is not None:
        vlb.x[part] += center[0]
        vlb.y[part] += center[1]
        vlb.z[part] += center[2]


def copy_beam(beamTo, beamFrom, indarr, includeState=False, includeJspEsp=True):
    for f in ('x', 'y', 'z', 'a', 'b', 'c', 'path', 'E'):
        getattr(beamTo, f)[indarr] = getattr(beamFrom, f)[indarr]
    if includeState:
        beamTo.state[indarr] = beamFrom.state[indarr]
    if includeJspEsp:
        beamTo.Jss[indarr] = beamFrom.Jss[indarr]
        beamTo.Jpp[indarr] = beamFrom.Jpp[indarr]
        beamTo.Jsp[indarr] = beamFrom.Jsp[indarr]
        if hasattr(beamFrom, 'Es') and hasattr(beamTo, 'Es'):
            beamTo.Es[indarr] = beamFrom.Es[indarr]
            beamTo.Ep[indarr] = beamFrom.Ep[indarr]


def rotate_coherency_matrix(beam, indarr, roll):
    c = np.cos(roll)
    s = np.sin(roll)
    c2 = c**2
    s2 = s**2
    cs = c * s
    JssN = beam.Jss[indarr]*c2 + beam.Jpp[indarr]*s2 + 2*beam.Jsp[indarr].real*cs
    JppN = beam.Jss[indarr]*s2 + beam.Jpp[indarr]*c2 - 2*beam.Jsp[indarr].real*cs
    JspN = (beam.Jpp[indarr]-beam.Jss[indarr])*cs + beam.Jsp[indarr].real*(c2-s2) + beam.Jsp[indarr].imag*1j
    return JssN, JppN, JspN
