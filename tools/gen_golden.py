#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (numpy path) in this container.

    python tools/gen_golden.py [section ...]          # sections: kirchhoff undulator materials reflect screen

The reference lives at /root/reference (read-only) and needs two import stubs (tools/ref_stubs: colorama,
matplotlib) because those packages are not installed here.  Nothing in tests/, smoke()
or bench.py reads /root/reference; they replay the committed .npz files.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, '/root/reference')
sys.path.insert(0, os.path.join(HERE, 'ref_stubs'))

import numpy as np  # noqa: E402

import xrt.backends.raycing as raycing  # noqa: E402
import xrt.backends.raycing.sources as rs  # noqa: E402
import xrt.backends.raycing.oes as roe  # noqa: E402
import xrt.backends.raycing.materials as rm  # noqa: E402
import xrt.backends.raycing.screens as rsc  # noqa: E402
import xrt.backends.raycing.apertures as ra  # noqa: E402
import xrt.backends.raycing.waves as rw  # noqa: E402
from xrt.backends.raycing import physconsts as rpc  # noqa: E402

raycing._VERBOSITY_ = 0
GOLD = os.path.join(ROOT, 'tests', 'golden')
os.makedirs(GOLD, exist_ok=True)


def save(name, **kw):
    path = os.path.join(GOLD, name + '.npz')
    np.savez_compressed(path, **kw)
    print('wrote', path, os.path.getsize(path), 'bytes')


BEAM_FIELDS = ('x', 'y', 'z', 'a', 'b', 'c', 'path', 'E', 'Jss', 'Jpp', 'Jsp', 'state', 'Es', 'Ep',
               'theta', 's', 'phi', 'r')


def beam_dict(beam, prefix):
    d = {}
    for f in BEAM_FIELDS:
        if hasattr(beam, f):
            v = getattr(beam, f)
            if isinstance(v, np.ndarray):
                d[prefix + f] = v
    return d


# ------------------------------------------------------------------------------------------------ constants
def gen_constants():
    names = ['PI', 'PI2', 'SIE0', 'C', 'M0', 'SIM0', 'M0C2', 'HPLANCK', 'EV2ERG', 'EMC', 'FINE_STR', 'E2W',
             'E2WC', 'R0', 'AVOGADRO', 'CHeVcm', 'CH', 'CHBAR', 'K2B']
    save('constants', **{n: np.float64(getattr(rpc, n)) for n in names},
         zEps=np.float64(raycing.zEps), dt=np.float64(raycing.dt),
         maxIteration=np.int64(raycing.maxIteration), maxHalfSizeOfOE=np.float64(raycing.maxHalfSizeOfOE),
         maxDepthOfOE=np.float64(raycing.maxDepthOfOE))


# ------------------------------------------------------------------------------------------------ kirchhoff
def _kirchhoff_case(seed, N, M, E0, dist, halfIn, halfOut, tilt):
    rng = np.random.RandomState(seed)
    oeLocal = rs.Beam(N, withAmplitudes=True)
    oeLocal.x[:] = rng.uniform(-halfIn, halfIn, N)
    oeLocal.z[:] = rng.uniform(-halfIn, halfIn, N)
    oeLocal.y[:] = rng.normal(0, 1e-3, N) if tilt else 0.
    oeLocal.state[:] = 1
    oeLocal.state[rng.rand(N) < 0.07] = 2  # some samples are not good and must be skipped
    oeLocal.E[:] = E0 * (1 + 1e-4 * rng.uniform(-1, 1, N))
    oeLocal.Es[:] = rng.normal(size=N) + 1j * rng.normal(size=N)
    oeLocal.Ep[:] = 0.3 * (rng.normal(size=N) + 1j * rng.normal(size=N))
    a = rng.normal(0, 1e-4, N)
    c = rng.normal(0, 1e-4, N)
    oeLocal.a[:] = a
    oeLocal.c[:] = c
    oeLocal.b[:] = np.sqrt(1 - a**2 - c**2)
    good = oeLocal.state == 1
    if tilt:
        nvec = rng.normal(size=(3, N)) * 0.05 + np.array([[0.], [1.], [0.]])
        nvec /= np.linalg.norm(nvec, axis=0)
        n = [nvec[0][good], nvec[1][good], nvec[2][good]]
        nfull = nvec
        nl = (oeLocal.a * nvec[0] + oeLocal.b * nvec[1] + oeLocal.c * nvec[2])
    else:
        n = [0, 1, 0]  # aperture branch of waves.diffract (waves.py:687-690)
        nfull = np.array([[0.], [1.], [0.]]) * np.ones((3, N))
        nl = oeLocal.a * n[0] + oeLocal.b * n[1] + oeLocal.c * n[2]

    wave = rs.Beam(M, withAmplitudes=True)
    wave.xDiffr = rng.uniform(-halfOut, halfOut, M)
    wave.zDiffr = rng.uniform(-halfOut, halfOut, M)
    wave.yDiffr = dist + rng.uniform(-1, 1, M) * (1. if tilt else 0.)
    res = rw._diffraction_integral_conv(oeLocal, n, nl, wave, good)
    return dict(x=oeLocal.x, y=oeLocal.y, z=oeLocal.z, E=oeLocal.E, Es=oeLocal.Es, Ep=oeLocal.Ep,
                state=oeLocal.state, n=nfull, nl=nl,
                xDiffr=wave.xDiffr, yDiffr=wave.yDiffr, zDiffr=wave.zDiffr,
                out_Es=res[0], out_Ep=res[1], out_aE=res[2], out_bE=res[3], out_cE=res[4])


def gen_kirchhoff():
    # hard X-rays over 10 m: phases ~ 4.6e11 rad (ulp 6e-5 rad) -- the case that needs bit-exact r and k*r
    save('kirchhoff_9keV_10m', **_kirchhoff_case(11, 700, 96, 9000., 1e4, 0.05, 0.5, False))
    # soft X-rays over 1 m with per-sample normals (OE branch, waves.py:675-686)
    save('kirchhoff_280eV_1m', **_kirchhoff_case(12, 500, 64, 280., 1e3, 2.0, 1.0, True))



# ------------------------------------------------------------------------------------------------ undulator
def _undulator_source(**kw):
    """tests/speed/2_synchrotronSources_speed.py:136-155 base parameters (BASELINE.json configs[1])."""
    bl = raycing.BeamLine()
    kwargs = dict(eE=3., eI=0.5, period=30., n=40, K=1.45, eSigmaX=48.65, eSigmaZ=6.197, eEpsilonX=0.263,
                  eEpsilonZ=0.008, xPrimeMax=0.1, zPrimeMax=0.1, xPrimeMaxAutoReduce=False,
                  zPrimeMaxAutoReduce=False, targetOpenCL=None)
    kwargs.update(kw)
    return rs.Undulator(bl, **kwargs)


def _undulator_dump(name, src, nsamp, seed, eEspread_seed=None):
    src.reset() if src.needReset else None
    rng = np.random.RandomState(seed)
    w = rng.uniform(src.E_min, src.E_max, nsamp)
    th = rng.uniform(src.Theta_min, src.Theta_max, nsamp)
    ps = rng.uniform(src.Psi_min, src.Psi_max, nsamp)
    gamma0 = src.gamma
    if src.eEspread > 0:
        np.random.seed(eEspread_seed)
        gam = gamma0 + gamma0 * src.eEspread * np.random.normal(size=nsamp)  # synchr.py:2053-2059
        np.random.seed(eEspread_seed)
    else:
        gam = gamma0 * np.ones(nsamp)
    I, Es, Ep = src.build_I_map(w, th, ps)
    assert src.gamma == gamma0 or src.eEspread > 0
    save(name, L0=src.L0, Np=src.Np, Kx=src.Kx, Ky=src.Ky, eI=src.eI, gamma0=gamma0, phase=src.phase,
         quadm=src.quadm, gIntervals=src.gIntervals, eEspread=src.eEspread,
         tg=src.tg, ag=src.ag, sintg=src.sintg, costg=src.costg, sintgph=src.sintgph, costgph=src.costgph,
         dstep=src.dstep, E_min=src.E_min, E_max=src.E_max, Theta_min=src.Theta_min, Theta_max=src.Theta_max,
         Psi_min=src.Psi_min, Psi_max=src.Psi_max, E1=src.E1, dx=src.dx, dz=src.dz, dxprime=src.dxprime,
         dzprime=src.dzprime, w=w, theta=th, psi=ps, gamma=gam, I=I, Es=Es, Ep=Ep,
         taperVal=np.float64(np.nan if src._taperVal is None else src._taperVal),
         R0=np.float64(np.nan if src.R0 is None else src.R0))


def gen_undulator():
    # planar undulator of config 2, first harmonic +-0.5 %, grid found by the reference's convergence search
    probe = _undulator_source(eMin=1000., eMax=1001., gNodes=16)
    E1 = probe.E1
    src = _undulator_source(eMin=E1*0.995, eMax=E1*1.005)
    _undulator_dump('undulator_planar_h1', src, 1500, 21)
    print('converged quadm', src.quadm, 'gIntervals', src.gIntervals, 'E1', E1)
    # helical-ish undulator (Kx != 0, 90 deg phase), fixed grid, wide band incl. 3rd harmonic, energy spread
    src = _undulator_source(K=0, Kx=0.9, Ky=1.2, phaseDeg=90, eMin=E1*0.5, eMax=E1*3.5, gNodes=40, gIntervals=3,
                            eEspread=8e-4)
    _undulator_dump('undulator_helical_spread', src, 900, 22, eEspread_seed=5)


def gen_undulator_ext():
    """f3: tapered and near-field undulators (explicit sum over the periods, synchr.py:1947-1996)."""
    probe = _undulator_source(eMin=1000., eMax=1001., gNodes=16)
    E1 = probe.E1
    # planar, linear gap taper 0.5 mm over the device at a 10 mm gap
    src = _undulator_source(eMin=E1*0.97, eMax=E1*1.03, gNodes=24, gIntervals=2, taper=(0.5, 10.))
    _undulator_dump('undulator_taper', src, 240, 23)
    # elliptical with taper (exercises the Kx terms of the taper branch), 3rd harmonic in the band
    src = _undulator_source(K=0, Kx=0.7, Ky=1.1, phaseDeg=90, n=12, eMin=E1*0.8, eMax=E1*3.6, gNodes=20,
                            gIntervals=3, taper=(-0.3, 12.))
    _undulator_dump('undulator_taper_elliptical', src, 240, 24)
    # near field: screen 25 m from a planar device; xPrimeMax wide enough for the finite-distance terms to matter
    src = _undulator_source(eMin=E1*0.97, eMax=E1*1.03, gNodes=24, gIntervals=2, R0=25000.)
    _undulator_dump('undulator_nf', src, 240, 25)
    # near field, elliptical, short distance (5 m), energy spread
    src = _undulator_source(K=0, Kx=0.7, Ky=1.1, phaseDeg=90, n=12, eMin=E1*0.8, eMax=E1*3.6, gNodes=20,
                            gIntervals=3, R0=5000., eEspread=8e-4)
    _undulator_dump('undulator_nf_elliptical', src, 240, 26, eEspread_seed=6)

# ------------------------------------------------------------------------------------------------ reflect
def _source_beam(bl, nrays, seed, withAmplitudes=True, dx=0.1, dz=0.05, dxprime=1e-4, dzprime=5e-5,
                 energies=(8990, 9010), polarization='h', kill=0.03):
    np.random.seed(seed)
    src = rs.GeometricSource(bl, 'src', nrays=nrays, dx=dx, dz=dz, dxprime=dxprime, dzprime=dzprime,
                             distE='flat', energies=energies, polarization=polarization)
    beam = src.shine(withAmplitudes=withAmplitudes)
    rng = np.random.RandomState(seed + 1000)
    dead = rng.rand(nrays) < kill
    beam.state[dead] = -7  # rays absorbed upstream must pass through untouched
    if withAmplitudes:
        # GeometricSource gives real unit amplitudes: make them generic complex so that every term is exercised
        ph = np.exp(1j * rng.uniform(0, 2*np.pi, nrays))
        beam.Es *= ph
        beam.Ep = beam.Ep * ph + 0.2 * np.abs(beam.Es) * np.exp(1j * rng.uniform(0, 2*np.pi, nrays))
        beam.Jss[:] = np.abs(beam.Es)**2
        beam.Jpp[:] = np.abs(beam.Ep)**2
        beam.Jsp[:] = beam.Es * np.conj(beam.Ep)
    return beam


def _oe_params(oe):
    d = dict(center=np.array(oe.center, dtype=float), sinAzimuth=oe.bl.sinAzimuth, cosAzimuth=oe.bl.cosAzimuth,
             pitch=float(oe.pitch), roll=float(oe.roll), positionRoll=float(oe.positionRoll), yaw=float(oe.yaw),
             extraPitch=float(oe.extraPitch), extraRoll=float(oe.extraRoll), extraYaw=float(oe.extraYaw),
             rotationSequence=np.array(oe.rotationSequence), lostNum=np.int64(oe.lostNum),
             surfPhysX=np.array(oe.surfPhysX, dtype=float), surfPhysY=np.array(oe.surfPhysY, dtype=float),
             overEdge=np.array(str(getattr(oe, 'overEdge', 'yMax'))), dx=float(oe.dx))
    if oe.surfOptX is not None:
        d['surfOptX'] = np.array(oe.surfOptX, dtype=float)
    if oe.surfOptY is not None:
        d['surfOptY'] = np.array(oe.surfOptY, dtype=float)
    return d


def _capture_reflect(oe, beam):
    """Run oe.reflect and also record the beam-wide scalars the numpy code derives on the way."""
    gb, lb = oe.reflect(beam)
    d = {}
    d.update(beam_dict(beam, 'in_'))
    d.update(beam_dict(gb, 'gb_'))
    d.update(beam_dict(lb, 'lb_'))
    d['footprint'] = np.array(oe.footprint[-1])
    return d


def gen_reflect():
    # (1) config 1: flat Au mirror at 4 mrad, small footprint limits so that rays are lost/over/out as well
    bl = raycing.BeamLine(azimuth=0.3)  # non-zero azimuth exercises the Rz part of the frame change
    beam = _source_beam(bl, 3000, 31)
    Au = rm.Material('Au', rho=19.3)
    p = 1e4
    ctr = (p*np.sin(0.3), p*np.cos(0.3), 0)
    oe = roe.OE(bl, 'M1', center=ctr, pitch=4e-3, roll=2e-3, yaw=-1e-3, material=Au,
                limPhysX=[-0.25, 0.3], limPhysY=[-60, 50], limOptX=[-0.2, 0.25], limOptY=[-45, 40])
    d = _capture_reflect(oe, beam)
    d.update(_oe_params(oe))
    save('reflect_flat_Au', rho=19.3, **d)

    # (2) toroid Au mirror (config 2 geometry), beamline along y, p-polarised part present
    bl = raycing.BeamLine()
    beam = _source_beam(bl, 3000, 32, dx=0.5, dz=0.3, dxprime=2e-4, dzprime=1e-4, energies=(6000, 12000))
    oe = roe.ToroidMirror(bl, 'TM', center=(0, 25000, 0), pitch=3e-3, R=8e6, r=75., material=Au,
                          limPhysX=[-4, 4], limPhysY=[-500, 400])
    d = _capture_reflect(oe, beam)
    d.update(_oe_params(oe))
    save('reflect_toroid_Au', rho=19.3, R=float(oe.R), r=float(oe.r), **d)

    # (3) strongly curved toroid without material and without amplitudes: drives the Brent branch
    beam = _source_beam(bl, 2000, 33, withAmplitudes=False, dx=1.0, dz=0.5, dxprime=1e-3, dzprime=1e-3)
    oe = roe.ToroidMirror(bl, 'TM2', center=(0, 3000, 0), pitch=0.15, R=4000., r=8., positionRoll=np.pi/2,
                          limPhysX=[-7, 7], limPhysY=[-200, 200])
    d = _capture_reflect(oe, beam)
    d.update(_oe_params(oe))
    save('reflect_toroid_bare', R=float(oe.R), r=float(oe.r), **d)

    # (4) a round (elliptical) flat Au mirror: shape='round' (rays_good, base.py:1140-1156)
    bl = raycing.BeamLine(azimuth=0.3)
    beam = _source_beam(bl, 3000, 34)
    oe = roe.OE(bl, 'Mr', center=ctr, pitch=4e-3, roll=2e-3, yaw=-1e-3, material=Au, shape='round',
                limPhysX=[-0.25, 0.3], limPhysY=[-60, 50])
    d = _capture_reflect(oe, beam)
    d.update(_oe_params(oe))
    st = d['gb_state']
    print('round', {int(x): int((st == x).sum()) for x in np.unique(st)})
    save('reflect_flat_round', rho=19.3, **d)

# ------------------------------------------------------------------------------------------------ planes
def gen_planes():
    bl = raycing.BeamLine(azimuth=-0.2)
    beam = _source_beam(bl, 2500, 51, dx=0.4, dz=0.3, dxprime=3e-4, dzprime=2e-4)
    beam.state[::17] = 2
    beam.state[5::29] = 3
    beam.b[7] = 0.  # a ray parallel to the screen: path = inf -> lost
    beam.c[7] = np.sqrt(1 - beam.a[7]**2)
    d = 7000.
    scr = rsc.Screen(bl, 'scr', center=(d*np.sin(-0.2) + 0.3, d*np.cos(-0.2), 1.5), compressX=0.7)
    blo = scr.expose(beam)
    out = dict(center=np.array(scr.center, dtype=float), ex=np.array(scr.x, dtype=float),
               ey=np.array(scr.y, dtype=float), ez=np.array(scr.z, dtype=float), lostNum=np.int64(scr.lostNum),
               compressX=0.7, footprint=np.array(scr.footprint[-1]), sinAzimuth=bl.sinAzimuth,
               cosAzimuth=bl.cosAzimuth)
    out.update(beam_dict(beam, 'in_'))
    out.update(beam_dict(blo, 'out_'))
    save('expose_screen', **out)
    # tilted screen given by explicit x and z, only positive paths
    scr2 = rsc.Screen(bl, 'scr2', center=(0., -300., 0.), x=(1., 0., 0.2), z=(-0.2, 0., 1.))
    blo = scr2.expose(beam, onlyPositivePath=True)
    out = dict(center=np.array(scr2.center, dtype=float), ex=np.array(scr2.x, dtype=float),
               ey=np.array(scr2.y, dtype=float), ez=np.array(scr2.z, dtype=float), lostNum=np.int64(scr2.lostNum))
    out.update(beam_dict(beam, 'in_'))
    out.update(beam_dict(blo, 'out_'))
    save('expose_screen_tilted', **out)

    beam = _source_beam(bl, 2500, 52, dx=0.4, dz=0.3, dxprime=3e-4, dzprime=2e-4)
    slit = ra.RectangularAperture(bl, 'slit', center=(4000*np.sin(-0.2), 4000*np.cos(-0.2), 0),
                                  kind=('left', 'right', 'bottom', 'top'), opening=[-0.9, 1.1, -0.6, 0.5])
    beam_in = rs.Beam(copyFrom=beam)
    lo = slit.propagate(beam)
    out = dict(center=np.array(slit.center, dtype=float), ex=np.array(slit.xyz[0], dtype=float),
               ey=np.array(slit.xyz[1], dtype=float), ez=np.array(slit.xyz[2], dtype=float),
               lostNum=np.int64(slit.lostNum), kinds=np.array(list(slit.blades.keys())),
               opening=np.array(list(slit.blades.values()), dtype=float),
               spotLimits=np.array(slit.spotLimits, dtype=float), in_state_after=beam.state)
    out.update(beam_dict(beam_in, 'in_'))
    out.update(beam_dict(lo, 'out_'))
    save('propagate_slit', **out)

# ------------------------------------------------------------------------------------------------ DCM + amplitudes
def _dcm_params(dcm):
    d = _oe_params(dcm)
    for k in ('bragg', 'cryst1roll', 'cryst2roll', 'cryst2pitch', 'cryst2finePitch', 'cryst2perpTransl',
              'cryst2longTransl'):
        d[k] = float(getattr(dcm, k))
    d['surfPhysX2'] = np.array(dcm.surfPhysX2, dtype=float)
    d['surfPhysY2'] = np.array(dcm.surfPhysY2, dtype=float)
    d['alpha'] = float(dcm.alpha) if dcm.alpha else 0.
    return d


def gen_dcm():
    bl = raycing.BeamLine()
    si = rm.CrystalSi(hkl=(1, 1, 1))
    E0 = 9000.
    for name, alpha in (('dcm_si111', None), ('dcm_si111_asym', np.radians(5.))):
        beam = _source_beam(bl, 3000, 61, dx=0.2, dz=0.02, dxprime=5e-5, dzprime=2e-5, energies=(8995, 9005))
        theta = si.get_Bragg_angle(E0) - si.get_dtheta(E0, alpha)
        dcm = roe.DCM(bl, 'dcm', center=(0, 2e4, 0), bragg=float(theta), material=si, material2=si, alpha=alpha,
                      cryst2perpTransl=10., cryst2finePitch=2e-6, cryst1roll=1e-4,
                      limPhysX=[-50, 50], limPhysY=[-100, 100], limPhysX2=[-50, 50], limPhysY2=[-100, 200])
        gb, lo1, lo2 = dcm.double_reflect(beam)
        d = _dcm_params(dcm)
        d.update(beam_dict(beam, 'in_'))
        d.update(beam_dict(gb, 'gb_'))
        d.update(beam_dict(lo1, 'lo1_'))
        d.update(beam_dict(lo2, 'lo2_'))
        save(name, d=si.d, tK=si.tK, E0=E0, **d)
        print(name, 'theta', theta, {int(s): int((gb.state == s).sum()) for s in np.unique(gb.state)})


def gen_amplitudes():
    # a12: Au / Pt / SiO2-like mirrors on an (E, angle) grid incl. the critical-angle region
    rng = np.random.RandomState(71)
    n = 4000
    E = rng.uniform(2000, 30000, n)
    th = rng.uniform(5e-4, 2e-2, n)
    bdn = -np.sin(th)
    out = dict(E=E, beamInDotNormal=bdn)
    for name, mat in (('Au', rm.Material('Au', rho=19.3)), ('Pt', rm.Material('Pt', rho=21.45)),
                      ('SiO2', rm.Material(('Si', 'O'), quantities=(1, 2), rho=2.2)),
                      ('AuThin', rm.Material('Au', rho=19.3, kind='thin mirror', t=2e-5))):
        res = mat.get_amplitude(E, bdn.copy())
        out[name + '_rs'], out[name + '_rp'] = res[0], res[1]
    save('amplitude_mirrors', **out)
    # a13: Si(111), Si(333) thick-Bragg rocking curves, symmetric and asymmetric (test_materials.py:239-352 recipe)
    out = {}
    for hkl in ((1, 1, 1), (3, 3, 3)):
        si = rm.CrystalSi(hkl=hkl)
        for alpha in (0., np.radians(5.), np.radians(-5.)):
            E = np.ones(600) * 10000.
            E[300:] = rng.uniform(9000, 11000, 300)
            thetaB = si.get_Bragg_angle(E)
            dth = np.linspace(-40, 80, 600) * 1e-6
            theta = thetaB + dth
            g0 = np.sin(theta + alpha)
            gh = -np.sin(theta - alpha)
            hns0 = np.sin(theta)
            cs, cp = si.get_amplitude(E, g0.copy(), gh.copy(), hns0.copy())
            key = 'Si%d%d%d_a%+d' % (hkl + (int(round(np.degrees(alpha))),))
            out[key + '_E'] = E
            out[key + '_g0'] = g0
            out[key + '_gh'] = gh
            out[key + '_hns0'] = hns0
            out[key + '_rs'] = cs
            out[key + '_rp'] = cp
            out[key + '_d'] = si.d
    save('amplitude_crystals', **out)

# ------------------------------------------------------------------------------------------------ shine
def gen_shine():
    probe = _undulator_source(eMin=1000., eMax=1001., gNodes=16)
    E1 = probe.E1
    for name, uniform, nrays, seed, spread in (('shine_uniform', True, 2500, 81, 0.), ('shine_rejection', False, 1200, 82, 0.),
                                               ('shine_uniform_spread', True, 2000, 83, 1e-3)):
        src = _undulator_source(eMin=E1*0.995, eMax=E1*1.005, gNodes=32, gIntervals=2, nrays=nrays, eEspread=spread,
                                uniformRayDensity=uniform, center=(0.5, -2., 0.1), pitch=1e-5, yaw=-2e-5)
        src.bl.azimuth = 0.05
        np.random.seed(seed)
        beam = src.shine()
        d = dict(L0=src.L0, Np=src.Np, Kx=src.Kx, Ky=src.Ky, eI=src.eI, gamma=src.gamma, phase=src.phase,
                 quadm=src.quadm, gIntervals=src.gIntervals, E_min=src.E_min, E_max=src.E_max,
                 Theta_min=src.Theta_min, Theta_max=src.Theta_max, Psi_min=src.Psi_min, Psi_max=src.Psi_max,
                 dx=src.dx, dz=src.dz, dxprime=src.dxprime, dzprime=src.dzprime, center=np.array(src.center, dtype=float),
                 sinAzimuth=src.bl.sinAzimuth, cosAzimuth=src.bl.cosAzimuth, pitch=src.pitch, yaw=src.yaw,
                 seed=seed, nrays=nrays, uniform=uniform, eEspread=spread, E1=src.E1, Imax=src.Imax, xzE=src.xzE, fluxConst=src.fluxConst,
                 accepted=beam.accepted, acceptedE=beam.acceptedE, seeded=beam.seeded, seededI=beam.seededI,
                 sourceWeight=beam.sourceWeight)
        d.update(beam_dict(beam, 'out_'))
        save(name, **d)
        print(name, 'seeded', beam.seeded, 'rays', len(beam.x))

# ------------------------------------------------------------------------------------------------ KB mirrors
def gen_kb():
    bl = raycing.BeamLine()
    Au = rm.Material('Au', rho=19.3)
    for name, kw in (('reflect_ellipse_vfm', dict(pitch=4e-3, p=30000., q=2000., isCylindrical=True)),
                     ('reflect_ellipse_hfm', dict(pitch=4e-3, p=30500., q=1500., isCylindrical=True,
                                                  positionRoll=np.pi/2)),
                     ('reflect_ellipsoid', dict(pitch=5e-3, p=20000., q=5000., isCylindrical=False))):
        beam = _source_beam(bl, 2500, 91, dx=0.05, dz=0.03, dxprime=8e-6, dzprime=6e-6, energies=(8000, 10000))
        oe = roe.EllipticalMirrorParam(bl, 'kb', center=(0, kw['p'], 0), material=Au,
                                       limPhysX=[-3, 3], limPhysY=[-90, 80], **kw)
        d = _capture_reflect(oe, beam)
        d.update(_oe_params(oe))
        for k in ('ellipseA', 'ellipseB', 'y0', 'z0', 'cosGamma', 'sinGamma', 'p', 'q'):
            d[k] = float(getattr(oe, k))
        d['isCylindrical'] = bool(oe.isCylindrical)
        st = d['gb_state']
        print(name, {int(x): int((st == x).sum()) for x in np.unique(st)})
        save(name, rho=19.3, **d)

# ------------------------------------------------------------------------------------------------ wave mode
def gen_diffract():
    bl = raycing.BeamLine(azimuth=0.1)
    d1, d2 = 1e4, 2e4
    sa, ca = np.sin(0.1), np.cos(0.1)
    slit = ra.RectangularAperture(bl, 'slit', center=(d1*sa, d1*ca, 0), kind=('left', 'right', 'bottom', 'top'),
                                  opening=[-0.05, 0.05, -0.04, 0.04])
    scr = rsc.Screen(bl, 'scr', center=(d2*sa, d2*ca, 0))
    N = 3000
    rng = np.random.RandomState(101)
    oeLocal = rs.Beam(N, withAmplitudes=True)
    oeLocal.x[:] = rng.uniform(-0.05, 0.05, N)
    oeLocal.z[:] = rng.uniform(-0.04, 0.04, N)
    oeLocal.y[:] = 0.
    r = np.sqrt(oeLocal.x**2 + d1**2 + oeLocal.z**2)
    oeLocal.a[:], oeLocal.b[:], oeLocal.c[:] = oeLocal.x / r, d1 / r, oeLocal.z / r
    oeLocal.E[:] = 9000.
    k = oeLocal.E / rpc.CHBAR * 1e7
    oeLocal.Es[:] = np.exp(1j * k * r) / r * 1e4
    oeLocal.Ep[:] = 0.05 * np.exp(1j * (k * r + 0.3)) / r * 1e4
    oeLocal.Jss[:] = np.abs(oeLocal.Es)**2
    oeLocal.Jpp[:] = np.abs(oeLocal.Ep)**2
    oeLocal.Jsp[:] = oeLocal.Es * np.conj(oeLocal.Ep)
    oeLocal.state[:] = 1
    oeLocal.state[rng.rand(N) < 0.03] = 2
    xm = np.linspace(-0.4, 0.4, 48)
    zm = np.linspace(-0.3, 0.3, 40)
    wave = scr.prepare_wave(slit, xm, zm)
    pre = dict(xDiffr=wave.xDiffr.copy(), yDiffr=wave.yDiffr.copy(), zDiffr=wave.zDiffr.copy(),
               rDiffr=wave.rDiffr.copy(), wa=wave.a.copy(), wb=wave.b.copy(), wc=wave.c.copy(), dS=wave.dS,
               warea=wave.area)
    glo = rw.diffract(oeLocal, wave)
    d = dict(azimuth=0.1, slit_center=np.array(slit.center, dtype=float), scr_center=np.array(scr.center, dtype=float),
             opening=np.array([-0.05, 0.05, -0.04, 0.04]), xmesh=xm, zmesh=zm, area=oeLocal.area, **pre)
    d.update(beam_dict(oeLocal, 'in_'))
    d.update(beam_dict(wave, 'wave_'))
    d.update(beam_dict(glo, 'glo_'))
    for k2 in ('EsAcc', 'EpAcc', 'aEacc', 'bEacc', 'cEacc'):
        d['wave_' + k2] = getattr(wave, k2)
    d['beamReflSumJ'] = wave.beamReflSumJ
    d['beamReflSumJnl'] = wave.beamReflSumJnl
    save('diffract_slit_screen', **d)

# ------------------------------------------------------------------------------------------------ Brent branch
def gen_brent():
    """A source sitting 0.06 mm above a strongly pitched toroid: |dz| at the near bracket end is tiny and at the
    far end large, which flips find_intersection to _use_Brent_method (base.py:871)."""
    import io
    import contextlib
    bl = raycing.BeamLine()
    p = 0.15
    lo = np.array([0., -30., 30.**2 / 2. / 2000. + 0.06])
    cy = -(np.cos(p) * lo[1] - np.sin(p) * lo[2])
    cz = -(np.sin(p) * lo[1] + np.cos(p) * lo[2])
    beam = _source_beam(bl, 2000, 111, dx=0.3, dz=0.004, dxprime=2e-3, dzprime=1e-3, energies=(8000, 9000))
    oe = roe.ToroidMirror(bl, 'TMB', center=(0, cy, cz), pitch=p, R=2000., r=15., material=rm.Material('Pt', rho=21.45),
                          limPhysX=[-5, 5], limPhysY=[-100, 100])
    raycing._VERBOSITY_ = 20
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        d = _capture_reflect(oe, beam)
    raycing._VERBOSITY_ = 0
    print(buf.getvalue().strip().splitlines()[-1])
    d.update(_oe_params(oe))
    st = d['gb_state']
    print('brent', {int(x): int((st == x).sum()) for x in np.unique(st)})
    save('reflect_toroid_brent', rho=21.45, R=float(oe.R), r=float(oe.r), **d)

# ------------------------------------------------------------------------------------------------ wave from an OE
def gen_diffract_oe():
    """Rays traced to a KB mirror (elliptical cylinder, parametric) are the diffracting samples; the field is
    propagated to a screen in the focus with rw.diffract (config 5, last step)."""
    bl = raycing.BeamLine()
    Au = rm.Material('Au', rho=19.3)
    pitch, p_, q_ = 4e-3, 30000., 2000.
    beam = _source_beam(bl, 3000, 121, dx=0.02, dz=0.01, dxprime=5e-6, dzprime=4e-6, energies=(8999.9, 9000.1),
                        kill=0.)
    kb = roe.EllipticalMirrorParam(bl, 'kb', center=(0, p_, 0), pitch=pitch, p=p_, q=q_, isCylindrical=True,
                                   material=Au, limPhysX=[-5, 5], limPhysY=[-200, 200])
    gb, lb = kb.reflect(beam)
    assert (lb.state == 1).all()
    scr = rsc.Screen(bl, 'scr', center=(0, p_ + q_*np.cos(2*pitch), q_*np.sin(2*pitch)))
    xm = np.linspace(-0.08, 0.08, 40)
    zm = np.linspace(-0.004, 0.004, 36)
    wave = scr.prepare_wave(kb, xm, zm)
    pre = dict(xDiffr=wave.xDiffr.copy(), yDiffr=wave.yDiffr.copy(), zDiffr=wave.zDiffr.copy(), dS=wave.dS)
    glo = rw.diffract(lb, wave)
    d = dict(scr_center=np.array(scr.center, dtype=float), xmesh=xm, zmesh=zm, area=lb.area, **pre)
    d.update(_oe_params(kb))
    for k in ('ellipseA', 'ellipseB', 'y0', 'z0', 'cosGamma', 'sinGamma', 'p', 'q'):
        d[k] = float(getattr(kb, k))
    d.update(beam_dict(lb, 'in_'))
    d.update(beam_dict(wave, 'wave_'))
    d.update(beam_dict(glo, 'glo_'))
    for k2 in ('EsAcc', 'EpAcc', 'aEacc', 'bEacc', 'cEacc'):
        d['wave_' + k2] = getattr(wave, k2)
    save('diffract_kb_screen', **d)

# ------------------------------------------------------------------------------------------------ wave onto an OE
def _slit_field(N, seed, d1, hx, hz, E0=9000.):
    rng = np.random.RandomState(seed)
    oeLocal = rs.Beam(N, withAmplitudes=True)
    oeLocal.x[:] = rng.uniform(-hx, hx, N)
    oeLocal.z[:] = rng.uniform(-hz, hz, N)
    oeLocal.y[:] = 0.
    r = np.sqrt(oeLocal.x**2 + d1**2 + oeLocal.z**2)
    oeLocal.a[:], oeLocal.b[:], oeLocal.c[:] = oeLocal.x / r, d1 / r, oeLocal.z / r
    oeLocal.E[:] = E0
    k = oeLocal.E / rpc.CHBAR * 1e7
    oeLocal.Es[:] = np.exp(1j * k * r) / r * 1e4
    oeLocal.Ep[:] = 0.05 * np.exp(1j * (k * r + 0.3)) / r * 1e4
    oeLocal.Jss[:] = np.abs(oeLocal.Es)**2
    oeLocal.Jpp[:] = np.abs(oeLocal.Ep)**2
    oeLocal.Jsp[:] = oeLocal.Es * np.conj(oeLocal.Ep)
    oeLocal.state[:] = 1
    return oeLocal


def _wave_onto_oe(name, mirror, slit, oeLocal, nrays, scr, xm, zm, seed, extra):
    """The body of OE.propagate_wave (oes/reflect.py:405-449) without the auto-alignment: prepare_wave on the
    mirror, diffract slit -> mirror, reflect(noIntersectionSearch=True), then mirror -> screen."""
    np.random.seed(seed)
    xy = None
    if isinstance(nrays, int):
        st = np.random.get_state()
        xy = np.random.rand(nrays, 2)
        np.random.set_state(st)
    waveOnMirror = mirror.prepare_wave(slit, nrays, rw=rw)
    d = dict(extra)
    if xy is not None:
        d['xy'] = xy
    d.update(beam_dict(oeLocal, 'in_'))
    d.update({k: v.copy() for k, v in beam_dict(waveOnMirror, 'pw_').items()})
    for k in ('xDiffr', 'yDiffr', 'zDiffr', 'rDiffr'):
        d['pw_' + k] = getattr(waveOnMirror, k).copy()
    d['pw_area'], d['pw_areaNormal'], d['pw_dS'] = waveOnMirror.area, waveOnMirror.areaNormal, waveOnMirror.dS
    beamToMirror = rw.diffract(oeLocal, waveOnMirror)
    d['slit_area'] = oeLocal.area
    d.update(beam_dict(waveOnMirror, 'wave_'))
    d.update(beam_dict(beamToMirror, 'glo_'))
    for k2 in ('EsAcc', 'EpAcc', 'aEacc', 'bEacc', 'cEacc'):
        d['wave_' + k2] = getattr(waveOnMirror, k2)
    mglo, mloc = mirror.reflect(beamToMirror, noIntersectionSearch=True)
    d.update(beam_dict(mglo, 'mglo_'))
    d.update(beam_dict(mloc, 'mloc_'))
    print(name, 'states on mirror', {int(x): int((mloc.state == x).sum()) for x in np.unique(mloc.state)})
    wave2 = scr.prepare_wave(mirror, xm, zm)
    glo2 = rw.diffract(mloc, wave2)
    d['mirror_area'] = mloc.area
    d.update(beam_dict(wave2, 'wave2_'))
    d.update(beam_dict(glo2, 'glo2_'))
    for k2 in ('EsAcc', 'EpAcc', 'aEacc', 'bEacc', 'cEacc'):
        d['wave2_' + k2] = getattr(wave2, k2)
    d.update(_oe_params(mirror))
    save(name, **d)


def gen_wave_oe():
    bl = raycing.BeamLine()
    Au = rm.Material('Au', rho=19.3)
    d1 = 1e4
    slit = ra.RectangularAperture(bl, 'slit', center=(0, d1, 0), kind=('left', 'right', 'bottom', 'top'),
                                  opening=[-0.05, 0.05, -0.04, 0.04])
    oeLocal = _slit_field(1500, 131, d1, 0.05, 0.04)
    # (1) toroid, random samples
    pitch = 4e-3
    tm = roe.ToroidMirror(bl, 'tm', center=(0, d1 + 3000., 0), pitch=pitch, R=1.2e6, r=80., material=Au,
                          limPhysX=[-0.4, 0.4], limPhysY=[-60, 60])
    scr = rsc.Screen(bl, 'scr', center=(0, d1 + 3000. + 2000.*np.cos(2*pitch), 2000.*np.sin(2*pitch)))
    xm = np.linspace(-0.3, 0.3, 24)
    zm = np.linspace(-0.3, 0.3, 20)
    _wave_onto_oe('wave_slit_toroid_screen', tm, slit, oeLocal, 1200, scr, xm, zm, 5,
                  dict(slit_center=np.array(slit.center, dtype=float), opening=np.array([-0.05, 0.05, -0.04, 0.04]),
                       scr_center=np.array(scr.center, dtype=float), xmesh=xm, zmesh=zm, R=float(tm.R), r=float(tm.r),
                       rho=19.3))
    # (2) elliptical cylinder (parametric), mesh samples, rolled to deflect horizontally
    oeLocal = _slit_field(1500, 132, d1, 0.05, 0.04)
    p_, q_ = d1 + 3000., 1500.
    kb = roe.EllipticalMirrorParam(bl, 'kb', center=(0, p_, 0), pitch=pitch, p=p_, q=q_, isCylindrical=True,
                                   material=Au, limPhysX=[-0.3, 0.3], limPhysY=[-50, 50])
    scr2 = rsc.Screen(bl, 'scr2', center=(0, p_ + q_*np.cos(2*pitch), q_*np.sin(2*pitch)))
    xm = np.linspace(-0.3, 0.3, 20)
    zm = np.linspace(-0.01, 0.01, 24)
    extra = dict(slit_center=np.array(slit.center, dtype=float), opening=np.array([-0.05, 0.05, -0.04, 0.04]),
                 scr_center=np.array(scr2.center, dtype=float), xmesh=xm, zmesh=zm, rho=19.3, mesh=np.array([24, 40]))
    for k in ('ellipseA', 'ellipseB', 'y0', 'z0', 'cosGamma', 'sinGamma', 'p', 'q'):
        extra[k] = float(getattr(kb, k))
    _wave_onto_oe('wave_slit_kb_screen', kb, slit, oeLocal, (24, 40), scr2, xm, zm, 6, extra)

# ------------------------------------------------------------------------------------------------ plot histograms
def gen_hist():
    """raycing.get_output (the reference's) on a traced screen beam for several flux kinds / ray flags; the
    histogram calls of multipro (np.histogram, np.histogram2d) are applied to its outputs by oracle/histogram.py
    (multipro itself needs matplotlib, absent here)."""
    import types
    sys.path.insert(0, ROOT)
    from oracle import histogram as oh
    bl = raycing.BeamLine()
    beam = _source_beam(bl, 8000, 141, dx=0.5, dz=0.2, dxprime=2e-4, dzprime=1e-4, energies=(8900, 9100),
                        polarization=None)
    oe = roe.ToroidMirror(bl, 'TM', center=(0, 1e4, 0), pitch=4e-3, R=1e6, r=100., material=rm.Material('Au', rho=19.3),
                          limPhysX=[-1.5, 1.5], limPhysY=[-150, 150], limOptX=[-1.2, 1.2], limOptY=[-120, 120])
    gb, lb = oe.reflect(beam)
    scr = rsc.Screen(bl, 'scr', center=(0, 2e4, 1e4*np.tan(8e-3)))
    blo = scr.expose(gb)
    blo.sourceWeight = 1.7e-3
    d = dict(beam_dict(blo, 'b_'), sourceWeight=blo.sourceWeight)
    xl, zl, el = [-4., 4.], [-1.5, 1.5], [8900., 9100.]
    d.update(xlimits=np.array(xl), ylimits=np.array(zl), climits=np.array(el), bins=np.array([64, 48, 32]),
             xfactor=1., yfactor=1e3, cfactor=1.)
    cases = [('total', (1,)), ('s', (1, 2)), ('p', (1, 3)), ('+/-45', (1, 2, 3)), ('left-right', (1,)),
             ('power', (1, 2, 3, -1))]
    for n, (fk, flags) in enumerate(cases):
        plot = types.SimpleNamespace(
            beam='b', showAbsorbed=False, beamState=None, rayFlag=flags, beamC=None, fluxKind=fk,
            xaxis=types.SimpleNamespace(data=raycing.get_x, factor=1.),
            yaxis=types.SimpleNamespace(data=raycing.get_z, factor=1e3),
            caxis=types.SimpleNamespace(data=raycing.get_energy, factor=1., useCategory=False))
        sel = raycing.get_output(plot, {'b': blo})
        for k, v in zip(('x', 'y', 'intensity', 'flux', 'cData'), sel[:5]):
            d['c%d_%s' % (n, k)] = v
        d['c%d_counts' % n] = np.array([int(v) for v in sel[5:11]])
        h = oh.plot_histograms(None, oh.Axis('x', xl, 64), oh.Axis('z', [zl[0]*1e3, zl[1]*1e3], 48, 1e3),
                               oh.Axis('E', el, 32), fk, flags, selected=sel)
        for k, v in h.items():
            d['c%d_%s' % (n, k)] = v
    d['fluxKinds'] = np.array([c[0] for c in cases])
    d['rayFlags'] = np.array([','.join(str(f) for f in c[1]) for c in cases])
    save('plot_histograms', **d)

# ------------------------------------------------------------------------------------------------ multilayers
def gen_multilayer():
    W, Si, Mo, C = (rm.Material('W', rho=19.3), rm.Material('Si', rho=2.33), rm.Material('Mo', rho=10.28),
                    rm.Material('C', rho=2.26))
    stacks = {
        'wsi': dict(tLayer=W, tThickness=12., bLayer=Si, bThickness=22., nPairs=60, substrate=Si, idThickness=2.),
        'moc_graded': dict(tLayer=Mo, tThickness=25., bLayer=C, bThickness=35., nPairs=40, substrate=Si,
                           tThicknessLow=15., bThicknessLow=20., idThickness=3., power=2.),
    }
    d = {}
    rng = np.random.RandomState(151)
    for name, kw in stacks.items():
        ml = rm.Multilayer(**kw)
        E0 = 9000.
        thB = float(ml.get_Bragg_angle(E0) - ml.get_dtheta(E0))
        n = 3000
        theta = np.linspace(0.3*thB, 2.2*thB, n)
        E = E0 + rng.uniform(-30, 30, n)
        bdn = -np.sin(theta)
        rs, rp = ml.get_amplitude(E, bdn)
        d.update({name + '_E': E, name + '_bdn': bdn, name + '_rs': rs, name + '_rp': rp, name + '_dti': ml.dti,
                  name + '_dbi': ml.dbi, name + '_thB': thB, name + '_d': ml.d})
        print(name, 'peak |rs|^2', (abs(rs)**2).max(), 'thetaB', thB)
    save('multilayer_amplitudes', **d)
    # a flat multilayer mirror in a beam line
    bl = raycing.BeamLine()
    ml = rm.Multilayer(**stacks['wsi'])
    thB = float(ml.get_Bragg_angle(9000.) - ml.get_dtheta(9000.))
    beam = _source_beam(bl, 4000, 152, dx=0.2, dz=0.05, dxprime=1e-4, dzprime=2e-4, energies=(8950, 9050))
    oe = roe.OE(bl, 'mlm', center=(0, 1e4, 0), pitch=thB, material=ml, limPhysX=[-5, 5], limPhysY=[-80, 80])
    dd = _capture_reflect(oe, beam)
    dd.update(_oe_params(oe))
    st = dd['gb_state']
    print('multilayer mirror', {int(x): int((st == x).sum()) for x in np.unique(st)},
          'mean R', float((dd['gb_Jss'] + dd['gb_Jpp'])[st == 1].mean() / (dd['in_Jss'] + dd['in_Jpp'])[st == 1].mean()))
    save('reflect_multilayer_flat', thB=thB, **dd)

# ------------------------------------------------------------------------------------------------ thin crystals
THIN = {'bragg_thin': ('Bragg reflected', 0.005, 0.), 'bragg_t': ('Bragg transmitted', 0.05, 0.),
        'laue_r': ('Laue reflected', 0.1, np.pi/2), 'laue_t': ('Laue transmitted', 0.1, np.pi/2)}


def gen_thin_crystals():
    rng = np.random.RandomState(161)
    out = {}
    for key, (geom, t, alpha) in THIN.items():
        si = rm.CrystalSi(hkl=(1, 1, 1), geom=geom, t=t)
        E = np.ones(600) * 10000.
        E[300:] = rng.uniform(9500, 10500, 300)
        thetaB = si.get_Bragg_angle(E)
        dth = np.linspace(-60, 90, 600) * 1e-6
        theta = thetaB + dth
        a = alpha + (np.radians(3.) if 'laue' in key else 0.)
        # the sign convention of _reflect_local: the beam goes INTO the surface (beamInDotNormal < 0)
        g0 = -np.sin(theta + a)
        gh = np.sin(theta - a)
        hns0 = -np.sin(theta)
        cs, cp = si.get_amplitude(E, g0.copy(), gh.copy(), hns0.copy())
        for k, v in (('E', E), ('g0', g0), ('gh', gh), ('hns0', hns0), ('rs', cs), ('rp', cp)):
            out[key + '_' + k] = v
        out[key + '_d'] = si.d
        print(key, 'max |rs|^2', float((abs(cs)**2).max()), 'nan', int(np.isnan(cs).sum()))
    save('amplitude_crystals_thin', **out)
    bl = raycing.BeamLine()
    E0 = 9000.
    for key, (geom, t, alpha) in THIN.items():
        si = rm.CrystalSi(hkl=(1, 1, 1), geom=geom, t=t)
        beam = _source_beam(bl, 1500, 162, dx=0.2, dz=0.02, dxprime=5e-5, dzprime=2e-5, energies=(8998, 9002))
        thetaB = float(si.get_Bragg_angle(E0))
        if 'laue' in key:
            pitch, al = np.pi/2 - thetaB, alpha
        else:
            pitch, al = thetaB - float(si.get_dtheta(E0)), None
        oe = roe.OE(bl, 'xtal', center=(0, 2e4, 0), pitch=pitch, material=si, alpha=al,
                    limPhysX=[-20, 20], limPhysY=[-20, 20])
        dd = _capture_reflect(oe, beam)
        dd.update(_oe_params(oe))
        st = dd['gb_state']
        good = st == 1
        print(key, {int(x): int((st == x).sum()) for x in np.unique(st)}, 'mean T/R',
              float((dd['gb_Jss'] + dd['gb_Jpp'])[good].mean() / (dd['in_Jss'] + dd['in_Jpp'])[good].mean()))
        save('reflect_crystal_' + key, d=si.d, tK=si.tK, t=t, geom=np.array(geom), alpha=0. if al is None else al, **dd)


def gen_gratings():
    """f4: gratings (toWhere = 3, reflect.py:840-861)."""
    bl = raycing.BeamLine()
    # (1) plane VLS grating, Au coated, order -1, soft X-rays at 2 deg grazing, density along y
    beam = _source_beam(bl, 2500, 51, dx=0.3, dz=0.1, dxprime=2e-4, dzprime=1e-4, energies=(395., 405.))
    Au = rm.Material('Au', rho=19.3, kind='grating')
    gd = ['y', 1200., 1., 2e-4, -3e-7, 1.5e-9]
    oe = roe.OE(bl, 'PG', center=(0, 20000, 0), pitch=np.radians(2.), material=Au, gratingDensity=gd, order=-1,
                limPhysX=[-8, 8], limPhysY=[-40, 35])
    d = _capture_reflect(oe, beam)
    d.update(_oe_params(oe))
    save('reflect_grating_vls', rho=19.3, gratingDensity=np.array(gd[1:], dtype=float), gratingAxis=np.array('y'),
         order=np.int64(-1), **d)
    # (2) toroidal grating without coating (EmptyMaterial), grooves along x (sagittal), order +2, 'auto' kind
    beam = _source_beam(bl, 2500, 52, dx=0.3, dz=0.1, dxprime=2e-4, dzprime=1e-4, energies=(900., 1000.))
    gd = ['x', 600., 1., -1e-3]
    oe = roe.ToroidMirror(bl, 'TG', center=(0, 15000, 0), pitch=np.radians(1.5), R=5e5, r=400.,
                          gratingDensity=gd, order=2, limPhysX=[-6, 6], limPhysY=[-60, 60])
    d = _capture_reflect(oe, beam)
    d.update(_oe_params(oe))
    save('reflect_grating_toroid', R=float(oe.R), r=float(oe.r), gratingDensity=np.array(gd[1:], dtype=float),
         gratingAxis=np.array('x'), order=np.int64(2), **d)
    # (3) constant line density through the default local_g (0, -100, 0) with tabulated efficiencies
    beam = _source_beam(bl, 2000, 53, dx=0.3, dz=0.1, dxprime=2e-4, dzprime=1e-4, energies=(395., 405.))
    mat = rm.Material('Au', rho=19.3, kind='grating', efficiency=[(1, 0.3), (2, 0.05)])
    oe = roe.OE(bl, 'PG2', center=(0, 20000, 0), pitch=np.radians(2.), material=mat, order=1,
                limPhysX=[-8, 8], limPhysY=[-40, 35])
    d = _capture_reflect(oe, beam)
    d.update(_oe_params(oe))
    save('reflect_grating_efficiency', rho=19.3, order=np.int64(1), eff_orders=np.array([1, 2]),
         eff_values=np.array([0.3, 0.05]), **d)


def _bm_source(kind, **kw):
    bl = raycing.BeamLine()
    kwargs = dict(eE=3., eI=0.5, eSigmaX=48.65, eSigmaZ=6.197, eEpsilonX=0.263, eEpsilonZ=0.008,
                  eMin=2000., eMax=22000., xPrimeMax=0.5, zPrimeMax=0.2)
    if kind == 'bm':
        kwargs.update(B0=1.7)
    else:
        kwargs.update(K=12., period=80., n=10)
    kwargs.update(kw)
    return (rs.BendingMagnet if kind == 'bm' else rs.Wiggler)(bl, **kwargs)


def _bm_params(src):
    return dict(gamma=src.gamma, B=src.B, ro=src.ro, Np=float(src.Np), isMPW=src.isMPW,
                K=float(getattr(src, 'K', 0.)), L0=float(getattr(src, 'L0', 0.)), X0=float(getattr(src, 'X0', 0.)),
                eI=src.eI, eEspread=src.eEspread, E_min=src.E_min, E_max=src.E_max, Theta_min=src.Theta_min,
                Theta_max=src.Theta_max, Psi_min=src.Psi_min, Psi_max=src.Psi_max, dx=src.dx, dz=src.dz,
                dxprime=src.dxprime, dzprime=src.dzprime, distE=np.array(src.distE))


def gen_bm():
    """f3: BendingMagnet / Wiggler (sources/synchr.py:69-609): build_I_map on seeded samples and seeded shine."""
    for kind in ('bm', 'wiggler'):
        src = _bm_source(kind)
        src.reset()
        rng = np.random.RandomState(61)
        n = 3000
        w = rng.uniform(src.E_min, src.E_max, n)
        th = rng.uniform(src.Theta_min, src.Theta_max, n)
        ps = rng.uniform(src.Psi_min, src.Psi_max, n)
        w[:40] = np.geomspace(5., 3e5, 40)  # eta from ~1e-4 to the underflow of K_nu
        I, Es, Ep = src.build_I_map(w, th, ps)
        save('%s_imap' % kind, w=w, theta=th, psi=ps, I=I, Es=Es, Ep=Ep, **_bm_params(src))
        for mode, uniform, nrays, seed, spread in (('uniform', True, 2500, 62, 0.), ('rejection', False, 1200, 63, 1e-3)):
            src = _bm_source(kind, nrays=nrays, uniformRayDensity=uniform, eEspread=spread,
                             center=(0.5, -2., 0.1), pitch=1e-5, yaw=-2e-5, eMin=5000., eMax=9000.)
            src.bl.azimuth = 0.05
            np.random.seed(seed)
            beam = src.shine()
            d = _bm_params(src)
            d.update(center=np.array(src.center, dtype=float), sinAzimuth=src.bl.sinAzimuth,
                     cosAzimuth=src.bl.cosAzimuth, pitch=src.pitch, yaw=src.yaw, seed=seed, nrays=nrays,
                     uniform=uniform, Imax=src.Imax, xzE=src.xzE, fluxConst=src.fluxConst, accepted=beam.accepted,
                     acceptedE=beam.acceptedE, seeded=beam.seeded, seededI=beam.seededI,
                     sourceWeight=beam.sourceWeight)
            d.update(beam_dict(beam, 'out_'))
            save('%s_shine_%s' % (kind, mode), **d)
            print(kind, mode, 'seeded', beam.seeded, 'rays', len(beam.x))


def _capture_refract(oe, beam, multiple=False):
    gb, l1, l2 = oe.multiple_refract(beam) if multiple else oe.double_refract(beam)
    d = {}
    d.update(beam_dict(beam, 'in_'))
    d.update(beam_dict(gb, 'gb_'))
    d.update(beam_dict(l1, 'lo1_'))
    d.update(beam_dict(l2, 'lo2_'))
    d.update(_oe_params(oe))
    d.update(t=float(oe.t), wedgeAngle=float(oe.wedgeAngle), surfPhysX2=np.array(oe.surfPhysX2, dtype=float),
             surfPhysY2=np.array(oe.surfPhysY2, dtype=float), rho=float(oe.material.rho))
    return d


def gen_refractive():
    """f4: plates and parabolic lenses (oes/refractive.py), refraction + transmittivity + absorption."""
    bl = raycing.BeamLine(azimuth=0.1)
    ctr = (1e4*np.sin(0.1), 1e4*np.cos(0.1), 0)
    # (1) diamond window, 0.3 mm, tilted 50 deg with a 2 mrad wedge
    beam = _source_beam(bl, 2500, 71, dx=0.4, dz=0.3, dxprime=2e-5, dzprime=2e-5, energies=(7000, 11000))
    C = rm.Material('C', rho=3.52, kind='plate')
    pl = roe.Plate(bl, 'W', center=ctr, pitch=np.radians(50.), t=0.3, wedgeAngle=2e-3, material=C,
                   limPhysX=[-0.9, 1.1], limPhysY=[-1.2, 1.0])
    save('refract_plate', **_capture_refract(pl, beam))
    # (2) single Be lens, paraboloid, zmax reached by part of the beam
    bl = raycing.BeamLine()
    beam = _source_beam(bl, 2500, 72, dx=0.35, dz=0.35, dxprime=1e-5, dzprime=1e-5, energies=(8500, 9500))
    Be = rm.Material('Be', rho=1.848, kind='lens')
    ln = roe.ParaboloidFlatLens(bl, 'L1', center=(0, 10000, 0), pitch=np.pi/2, t=0.05, material=Be, focus=0.2,
                                zmax=0.6, limPhysX=[-1, 1], limPhysY=[-1, 1])
    save('refract_lens_single', focus=0.2, zmax=0.6, nCRL=np.int64(1), **_capture_refract(ln, beam))
    # (3) CRL: 5 double paraboloid lenses through multiple_refract
    beam = _source_beam(bl, 2500, 73, dx=0.2, dz=0.2, dxprime=1e-5, dzprime=1e-5, energies=(8500, 9500))
    ln = roe.DoubleParaboloidLens(bl, 'CRL', center=(0, 10000, 0), pitch=np.pi/2, t=0.03, material=Be, focus=0.1,
                                  zmax=0.5, nCRL=5, limPhysX=[-0.6, 0.6], limPhysY=[-0.6, 0.6])
    save('refract_crl_double', focus=0.1, zmax=0.5, nCRL=np.int64(5), **_capture_refract(ln, beam, True))
    # (4) 1-D focusing: 3 parabolic cylinder lenses
    beam = _source_beam(bl, 2000, 74, dx=0.2, dz=0.2, dxprime=1e-5, dzprime=1e-5, energies=(8500, 9500))
    ln = roe.ParabolicCylinderFlatLens(bl, 'CL', center=(0, 10000, 0), pitch=np.pi/2, t=0.04,
                                       material=Be, focus=0.15, zmax=0.4, nCRL=3, limPhysX=[-0.5, 0.5],
                                       limPhysY=[-0.5, 0.5])
    save('refract_crl_cylinder', focus=0.15, zmax=0.4, nCRL=np.int64(3), **_capture_refract(ln, beam, True))


def _geom_case(name, seed, withAmplitudes, azimuth=0., **kw):
    bl = raycing.BeamLine(azimuth=azimuth)
    np.random.seed(seed)
    src = rs.GeometricSource(bl, 'gs', **kw)
    beam = src.shine(withAmplitudes=withAmplitudes)
    d = beam_dict(beam, 'out_')
    d.update(seed=seed, withAmplitudes=withAmplitudes, sinAzimuth=bl.sinAzimuth, cosAzimuth=bl.cosAzimuth)
    for k in ('sourceWeight', 'seeded', 'seededI', 'accepted', 'acceptedE'):
        if hasattr(beam, k):
            d[k] = getattr(beam, k)
    save(name, **d)


GEOM_CASES = {
    # config 1 of BASELINE.json
    'geom_config1': dict(seed=1234, withAmplitudes=True, kw=dict(
        nrays=4000, dx=0.1, dz=0.05, dxprime=1e-4, dzprime=5e-5, distE='flat', energies=(8990, 9010),
        polarization='h')),
    # uniformRayDensity with cut limits, flat x, annular divergence, weighted lines, circular polarisation, canted
    'geom_uniform_annulus': dict(seed=77, withAmplitudes=True, azimuth=0.2, kw=dict(
        nrays=3000, center=(1., -2., 0.5), distx='flat', dx=(-0.3, 0.5), distz='normal', dz=(0.02, 0.07),
        disty='flat', dy=2., distxprime='annulus', dxprime=(1e-4, 3e-4), distzprime='annulus', dzprime=(0.3, 2.5),
        distE='lines', energies=(8000., 9000., 12000.), energyWeights=(0.2, 0.5, 0.3), polarization='r',
        uniformRayDensity=True, pitch=1e-3, roll=0.3, yaw=-2e-3)),
    # unpolarised, normal energy band, everything normal incl. y, absolute flux
    'geom_unpolarized_flux': dict(seed=78, withAmplitudes=True, kw=dict(
        nrays=3000, disty='normal', dy=5., distE='normal', energies=(9000., 3.), polarization=None,
        totalFlux=1e12)),
    # no amplitudes, custom coherency matrix, unweighted lines, point source in z
    'geom_custom_pol': dict(seed=79, withAmplitudes=False, kw=dict(
        nrays=3000, distz=None, dz=0, distE='lines', energies=(5000., 7000.), polarization=(0.7, 0.3, 0.1, -0.2),
        dxprime=2e-3, dzprime=1.5e-3)),
}


def gen_geometric():
    """GeometricSource.shine (sources/geoms.py:193-534): the input generator of BASELINE.json configs 1 and 3."""
    for name, c in GEOM_CASES.items():
        _geom_case(name, c['seed'], c['withAmplitudes'], c.get('azimuth', 0.), **c['kw'])


def gen_mesh():
    """SourceBase.intensities_on_mesh (sybase.py:740-928): the flux-map use of the undulator kernel."""
    probe = _undulator_source(eMin=1000., eMax=1001., gNodes=16)
    E1 = probe.E1
    # (1) planar, harmonics 1 and 3 separated, energy spread (5 samples), emittance convolution
    src = _undulator_source(eMin=E1*0.9, eMax=E1*3.2, gNodes=40, gIntervals=2, eEspread=8e-4, distE='BW')
    energy = np.linspace(E1*0.9, E1*3.2, 14)
    theta = np.linspace(-1, 1, 21) * 8e-5
    psi = np.linspace(-1, 1, 15) * 6e-5
    harm = [1, 3]
    ss = src.intensities_on_mesh(energy, theta, psi, harm, eSpreadSigmas=3., eSpreadNSamples=5)
    save('mesh_planar_harmonics', energy=energy, theta=theta, psi=psi, harmonic=np.array(harm), s0=ss[0], s1=ss[1],
         s2=ss[2], s3=ss[3], L0=src.L0, Np=src.Np, Kx=src.Kx, Ky=src.Ky, eI=src.eI, gamma0=src.gamma,
         phase=src.phase, quadm=src.quadm, gIntervals=src.gIntervals, eEspread=src.eEspread, dxprime=src.dxprime,
         dzprime=src.dzprime, eSpreadSigmas=3., eSpreadNSamples=5, distE=np.array('BW'))
    # (2) elliptical, zero emittance, no spread, no harmonic axis
    src = _undulator_source(K=0, Kx=0.9, Ky=1.2, phaseDeg=90, eMin=E1*0.5, eMax=E1*1.5, gNodes=40, gIntervals=3,
                            eEpsilonX=0., eEpsilonZ=0., eSigmaX=0., eSigmaZ=0.)
    energy = np.linspace(E1*0.5, E1*1.5, 9)
    theta = np.linspace(-1, 1, 17) * 1e-4
    psi = np.linspace(-1, 1, 13) * 1e-4
    ss = src.intensities_on_mesh(energy, theta, psi)
    save('mesh_elliptical', energy=energy, theta=theta, psi=psi, s0=ss[0], s1=ss[1], s2=ss[2], s3=ss[3], L0=src.L0,
         Np=src.Np, Kx=src.Kx, Ky=src.Ky, eI=src.eI, gamma0=src.gamma, phase=src.phase, quadm=src.quadm,
         gIntervals=src.gIntervals, eEspread=src.eEspread, dxprime=src.dxprime, dzprime=src.dzprime,
         eSpreadSigmas=3.5, eSpreadNSamples=36, distE=np.array('eV'))


def gen_config1():
    """BASELINE.json configs[0] exactly as SURVEY.md 8d states it: np.random.seed(1234); GeometricSource(1e5 rays)
    -> flat Au mirror at 4 mrad -> Screen.  The fixture keeps every 50th ray of the three beams plus whole-beam sums
    (the full beams are 3 x 13 MB); the oracle reproduces the complete beams from the same seed."""
    np.random.seed(1234)
    bl = raycing.BeamLine()
    src = rs.GeometricSource(bl, 'src', nrays=100000, dx=0.1, dz=0.05, dxprime=1e-4, dzprime=5e-5, distE='flat',
                             energies=(8990, 9010), polarization='h')
    pitch = 4e-3
    oe = roe.OE(bl, 'M1', center=(0, 1e4, 0), pitch=pitch, material=rm.Material('Au', rho=19.3),
                limPhysX=[-20, 20], limPhysY=[-300, 300])
    scr = rsc.Screen(bl, 'scr', center=(0, 2e4, 1e4*np.tan(2*pitch)))
    beam = src.shine(withAmplitudes=True)
    gb, lb = oe.reflect(beam)
    blo = scr.expose(gb)
    d = dict(stride=50, nrays=100000, seed=1234, pitch=pitch, lostNum=np.int64(oe.lostNum),
             scr_lostNum=np.int64(scr.lostNum), footprint=np.array(oe.footprint[-1]))
    for pre, b in (('src_', beam), ('gb_', gb), ('lb_', lb), ('scr_', blo)):
        for f in BEAM_FIELDS:
            if hasattr(b, f) and isinstance(getattr(b, f), np.ndarray):
                v = getattr(b, f)
                d[pre + f] = v[::50]
                d[pre + f + '_sum'] = v.sum()
    save('config1_chain', **d)


def _custom_field(ncol):
    L0, Np = 30., 10
    z = np.linspace(-0.5*L0*Np - 30, 0.5*L0*Np + 30, 2401)
    env = 0.5*(np.tanh((z + 0.5*L0*Np)/8.) - np.tanh((z - 0.5*L0*Np)/8.))
    By = 0.5*np.sin(2*np.pi*z/L0)*env
    Bx = 0.2*np.cos(2*np.pi*z/L0)*env
    cols = {2: (z, By), 3: (z, Bx, By), 4: (z, Bx, By, 0.01*env)}[ncol]
    return np.vstack(cols).T


def gen_field_source():
    """f3: SourceFromField (sources/synchr.py:612-1347): tabulated field -> RK4 trajectory -> radiation integral."""
    for name, ncol, R0, spread, seed in (('field_source_far', 4, None, 0., 91), ('field_source_near', 2, 20000., 8e-4, 92)):
        bl = raycing.BeamLine()
        field = _custom_field(ncol)
        src = rs.SourceFromField(bl, eE=3., eI=0.5, eEpsilonX=0.263, eEpsilonZ=0.008, eSigmaX=48.65, eSigmaZ=6.197,
                                 xPrimeMax=0.1, zPrimeMax=0.1, eMin=1380., eMax=1520., customField=field, gNodes=60,
                                 gIntervals=12, R0=R0, eEspread=spread, targetOpenCL=None)
        src.deviceLength = src._fwhm(field[:, 0], np.abs(field[:, 1:]).max(axis=1))  # what read_custom_field sets
        src.reset() if src.needReset else None
        rng = np.random.RandomState(seed)
        n = 300
        w = rng.uniform(src.E_min, src.E_max, n)
        th = rng.uniform(src.Theta_min, src.Theta_max, n)
        ps = rng.uniform(src.Psi_min, src.Psi_max, n)
        gamma0 = src.gamma
        dg = gamma0 * spread * rng.normal(size=n) if spread > 0 else None
        I, Es, Ep = src.build_I_map(w, th, ps, None, dg)
        save(name, field=field, gamma0=gamma0, eI=src.eI, quadm=src.quadm, gIntervals=src.gIntervals, tg=src.tg,
             ag=src.ag, dstep=src.dstep, R0=np.float64(np.nan if R0 is None else R0), eEspread=spread,
             deviceLength=src.deviceLength, E_min=src.E_min, E_max=src.E_max, Theta_min=src.Theta_min,
             Theta_max=src.Theta_max, Psi_min=src.Psi_min, Psi_max=src.Psi_max, w=w, theta=th, psi=ps,
             gamma=gamma0 + (dg if dg is not None else 0.) * np.ones(n), I=I, Es=Es, Ep=Ep,
             beta_x=src.beta[0], beta_y=src.beta[1], traj_x=src.trajectory[0], traj_y=src.trajectory[1],
             traj_z=src.trajectory[2])


def gen_spliced():
    """Curved elements the product compiles from spliced CUDA texts (xrt_b200/oes_spliced.py): BentFlatMirror,
    ConicalMirror (oes/__init__.py:240-306, :589-636), JohannCylinder, JohanssonCylinder, JohannToroid
    (oes/bragg.py:104-269), the bent crystals with and without asymmetry angle."""
    bl = raycing.BeamLine(azimuth=0.1)
    Au = rm.Material('Au', rho=19.3)
    p = 2e4
    ctr = (p*np.sin(0.1), p*np.cos(0.1), 0)
    beam = _source_beam(bl, 2500, 201, dx=0.3, dz=0.1, dxprime=1.2e-4, dzprime=6e-5, energies=(7000, 11000))
    oe = roe.BentFlatMirror(bl, 'VCM', center=ctr, pitch=3e-3, R=4.2e6, material=Au, limPhysX=[-2, 2.5],
                            limPhysY=[-350, 300], limOptY=[-300, 280])
    d = _capture_reflect(oe, beam)
    d.update(_oe_params(oe))
    save('reflect_spliced_bentflat', rho=19.3, R=float(oe.R), **d)

    beam = _source_beam(bl, 2500, 202, dx=0.3, dz=0.1, dxprime=1.2e-4, dzprime=6e-5, energies=(7000, 11000))
    oe = roe.ConicalMirror(bl, 'cone', center=ctr, pitch=4e-3, roll=1e-3, L0=1500., theta=4e-3, material=Au,
                           limPhysX=[-2, 2.5], limPhysY=[-400, 380])
    d = _capture_reflect(oe, beam)
    d.update(_oe_params(oe))
    save('reflect_spliced_conical', rho=19.3, L0=float(oe.L0), theta=float(oe.theta), **d)

    # parametric: collimating paraboloid (p given) and focusing parabolic cylinder (q given, rolled to the side)
    bl = raycing.BeamLine()
    for name, kw, okw in (('parab_p', dict(p=20000.), dict(pitch=4e-3)),
                          ('parab_q_cyl', dict(q=8000., isCylindrical=True), dict(pitch=3e-3, positionRoll=np.pi/2))):
        beam = _source_beam(bl, 2500, 205, dx=0.3, dz=0.1, dxprime=1.2e-4, dzprime=6e-5, energies=(7000, 11000))
        oe = roe.ParabolicalMirrorParam(bl, name, center=(0, 20000., 0), material=Au, limPhysX=[-3, 2.5],
                                        limPhysY=[-400, 380], **dict(kw, **okw))
        d = _capture_reflect(oe, beam)
        d.update(_oe_params(oe))
        st = d['gb_state']
        print(name, {int(x): int((st == x).sum()) for x in np.unique(st)})
        save('reflect_spliced_' + name, rho=19.3, p=np.float64(np.nan if oe.p is None else oe.p),
             q=np.float64(np.nan if oe.q is None else oe.q), isCylindrical=bool(oe.isCylindrical),
             parabParam=float(oe.parabParam), y0=float(oe.y0), z0=float(oe.z0), cosGamma=float(oe.cosGamma),
             sinGamma=float(oe.sinGamma), **d)

    bl = raycing.BeamLine()
    si = rm.CrystalSi(hkl=(1, 1, 1))
    E0 = 9000.
    thetaB = float(si.get_Bragg_angle(E0))
    pdist = 500.
    Rm = pdist / np.sin(thetaB)  # the source sits on the Rowland circle (diameter Rm)
    cases = (('johann_circ', roe.JohannCylinder, dict(Rm=Rm, crossSection='circular'), None),
             ('johann_parab_alpha', roe.JohannCylinder, dict(Rm=Rm, crossSection='parabolic'), np.radians(4.)),
             ('johansson', roe.JohanssonCylinder, dict(Rm=Rm), None),
             ('johansson_alpha', roe.JohanssonCylinder, dict(Rm=Rm), np.radians(-3.)),
             ('johann_toroid', roe.JohannToroid, dict(Rm=Rm, Rs=Rm*np.sin(thetaB)**2), None),
             ('johann_toroid_alpha', roe.JohannToroid, dict(Rm=Rm, Rs=Rm*np.sin(thetaB)**2), np.radians(5.)))
    for i, (name, cls, kw, alpha) in enumerate(cases):
        beam = _source_beam(bl, 1500, 210 + i, dx=0.02, dz=0.02, dxprime=8e-3, dzprime=1.5e-2, energies=(8998, 9002))
        pitch = thetaB - float(si.get_dtheta(E0, alpha)) + (alpha or 0.)
        oe = cls(bl, name, center=(0, pdist, 0), pitch=pitch, material=si, alpha=alpha, limPhysX=[-4, 4],
                 limPhysY=[-30, 30], **kw)
        d = _capture_reflect(oe, beam)
        d.update(_oe_params(oe))
        st = d['gb_state']
        good = st == 1
        print(name, {int(x): int((st == x).sum()) for x in np.unique(st)}, 'mean R',
              float((d['gb_Jss'] + d['gb_Jpp'])[good].mean() / (d['in_Jss'] + d['in_Jpp'])[good].mean()))
        save('reflect_spliced_' + name, d=si.d, tK=si.tK, alpha=0. if alpha is None else alpha,
             Rm=float(oe.Rm), Rs=float(getattr(oe, 'Rs', 0.)), **d)


def gen_multi():
    """OE.multiple_reflect (oes/reflect.py:165-264): whispering-gallery passes along a strongly bent mirror -- the
    meridional cylinder BentFlatMirror and a ToroidMirror -- up to 8 reflections, rays lost / over on the way."""
    bl = raycing.BeamLine()
    Au = rm.Material('Au', rho=19.3)
    for name, cls, kw in (('bentflat', roe.BentFlatMirror, dict(R=2000.)),
                          ('toroid', roe.ToroidMirror, dict(R=2000., r=60.))):
        beam = _source_beam(bl, 600, 301, dx=0.05, dz=0.02, dxprime=1e-4, dzprime=5e-4, energies=(8000, 9000))
        oe = cls(bl, 'wg', center=(0, 1000, 4.9 if name == 'bentflat' else 2.4), pitch=0.05, material=Au,
                 limPhysX=[-5, 5], limPhysY=[-100, 100], **kw)
        gb, lbN = oe.multiple_reflect(beam, maxReflections=8)
        d = {}
        d.update(beam_dict(beam, 'in_'))
        d.update(beam_dict(gb, 'gb_'))
        d.update(beam_dict(lbN, 'lbN_'))
        d['gb_nRefl'] = gb.nRefl
        d['lbN_nRefl'] = lbN.nRefl
        d['footprints'] = np.array(oe.footprint)
        d.update(_oe_params(oe))
        print(name, 'passes', len(oe.footprint), 'nRefl', dict(zip(*np.unique(gb.nRefl, return_counts=True))),
              'state', dict(zip(*np.unique(gb.state, return_counts=True))))
        save('multi_reflect_' + name, rho=19.3, R=float(oe.R), r=float(getattr(oe, 'r', 0.)), **d)
    # the same passes with needElevationMap=True (reflect.py:212-216, :651-659; base.py:1279-1289): the highest point of
    # a ray above the surface between two impacts and its distance, per pass.  A second beam line with the same two
    # elements in the same order (lostNum is the element's ordinal), so that the files above stay byte-identical.
    bl = raycing.BeamLine()
    for name, cls, kw in (('bentflat', roe.BentFlatMirror, dict(R=2000.)),
                          ('toroid', roe.ToroidMirror, dict(R=2000., r=60.))):
        beam = _source_beam(bl, 600, 301, dx=0.05, dz=0.02, dxprime=1e-4, dzprime=5e-4, energies=(8000, 9000))
        oe = cls(bl, 'wg', center=(0, 1000, 4.9 if name == 'bentflat' else 2.4), pitch=0.05, material=Au,
                 limPhysX=[-5, 5], limPhysY=[-100, 100], **kw)
        gbe, lbNe = oe.multiple_reflect(beam, maxReflections=8, needElevationMap=True)
        ref = np.load(os.path.join(GOLD, 'multi_reflect_' + name + '.npz'))
        assert np.array_equal(gbe.x, ref['gb_x']) and np.array_equal(lbNe.x, ref['lbN_x'])
        assert np.array_equal(gbe.state, ref['gb_state'])
        e = {}
        for pre, b in (('gb_', gbe), ('lbN_', lbNe)):
            for f in ('elevationD', 'elevationX', 'elevationY', 'elevationZ'):
                e[pre + f] = getattr(b, f)
        print(name, 'elevation D max', float(lbNe.elevationD.max()), 'rays with a map', int((lbNe.elevationD > -1).sum()))
        save('multi_reflect_' + name + '_elevation', **e)


def gen_shine_wave():
    """IntegratedSource.shine(wave=...) (sources/sybase.py:1513-1527, :1588-1612, :1651-1703, :1786-1796): the source
    illuminates the receiving points of a wave prepared by Screen.prepare_wave -- filament beam (one electron, with
    energy spread) and the full emittance (per-sample shifts)."""
    probe = _undulator_source(eMin=1000., eMax=1001., gNodes=16)
    E1 = probe.E1
    for name, filament, seed, spread in (('shine_wave_filament', True, 91, 8e-4), ('shine_wave_full', False, 92, 0.)):
        src = _undulator_source(eMin=E1*0.999, eMax=E1*1.001, gNodes=32, gIntervals=2, nrays=1000, eEspread=spread,
                                uniformRayDensity=True, filamentBeam=filament, center=(0.3, -1., 0.05))
        src.bl.azimuth = 0.02
        scr = rsc.Screen(src.bl, 'scr', center=(25000.*np.sin(0.02), 25000.*np.cos(0.02), 0.2))
        xm = np.linspace(-2., 2., 31)
        zm = np.linspace(-1.5, 1.5, 25)
        wave = scr.prepare_wave(src, xm, zm)
        d = dict(xDiffr=wave.xDiffr.copy(), yDiffr=wave.yDiffr.copy(), zDiffr=wave.zDiffr.copy(), area=wave.area,
                 dS=wave.dS, wave_in_x=wave.x.copy(), wave_in_y=wave.y.copy(), wave_in_z=wave.z.copy())
        # a filament source probes Imax with nrays random samples inside reset() (sybase.py:529-541): do that before
        # seeding, so that the seeded stream starts at shine, and hand the probed Imax to the oracle
        np.random.seed(seed + 1000)
        if src.needReset:
            src.reset()
        d['Imax0'] = src.Imax
        np.random.seed(seed)
        bor = src.shine(wave=wave)
        d.update(L0=src.L0, Np=src.Np, Kx=src.Kx, Ky=src.Ky, eI=src.eI, gamma=src.gamma, phase=src.phase,
                 quadm=src.quadm, gIntervals=src.gIntervals, E_min=src.E_min, E_max=src.E_max,
                 Theta_min=src.Theta_min, Theta_max=src.Theta_max, Psi_min=src.Psi_min, Psi_max=src.Psi_max,
                 dx=src.dx, dz=src.dz, dxprime=src.dxprime, dzprime=src.dzprime,
                 center=np.array(src.center, dtype=float), sinAzimuth=src.bl.sinAzimuth, cosAzimuth=src.bl.cosAzimuth,
                 seed=seed, filament=filament, eEspread=spread, E1=src.E1, Imax=src.Imax, xzE=src.xzE,
                 fluxConst=src.fluxConst, accepted=np.copy(bor.accepted), acceptedE=np.copy(bor.acceptedE),
                 seeded=np.copy(bor.seeded),  # copies: a later shine(accuBeam=bor) adds to these 0-d arrays in place
                 seededI=np.copy(bor.seededI), sourceWeight=np.copy(bor.sourceWeight), xm=xm, zm=zm,
                 screen_center=np.array(scr.center, dtype=float), rDiffr=wave.rDiffr)
        if filament:
            # the same electron illuminates a second wave: shine(wave=wave2, accuBeam=bor) (sybase.py:1553-1561)
            wave2 = scr.prepare_wave(src, xm[::2] + 0.3, zm[::2] - 0.2)
            d.update(w2_xDiffr=wave2.xDiffr.copy(), w2_yDiffr=wave2.yDiffr.copy(), w2_zDiffr=wave2.zDiffr.copy(),
                     w2_area=wave2.area)
            bor2 = src.shine(wave=wave2, accuBeam=bor)
            d.update(beam_dict(wave2, 'w2_'))
            d.update(w2_rDiffr=wave2.rDiffr, w2_seeded=bor2.seeded, w2_seededI=bor2.seededI,
                     w2_bor_x=bor2.x, w2_bor_a=bor2.a)
            d.update(filamentDX=bor.filamentDX, filamentDZ=bor.filamentDZ, filamentDtheta=bor.filamentDtheta,
                     filamentDpsi=bor.filamentDpsi, filamentDgamma=bor.filamentDgamma)
        d.update(beam_dict(bor, 'bor_'))
        d.update(beam_dict(wave, 'wave_'))
        save(name, **d)
        print(name, 'samples', len(wave.x), 'max |Es|', float(np.abs(wave.Es).max()))


def gen_round():
    """RoundAperture.propagate (apertures.py:769-825): a pipe and a round beam stop; RoundAperture.prepare_wave
    (:848-875) with seeded samples in the disk."""
    bl = raycing.BeamLine(azimuth=-0.2)
    for name, stop in (('propagate_round', False), ('propagate_round_stop', True)):
        beam = _source_beam(bl, 2500, 53, dx=0.4, dz=0.3, dxprime=3e-4, dzprime=2e-4)
        ap = ra.RoundAperture(bl, 'pipe', center=(4000*np.sin(-0.2), 4000*np.cos(-0.2), 0.1), r=1.05)
        ap.isBeamStop = stop
        beam_in = rs.Beam(copyFrom=beam)
        lo = ap.propagate(beam)
        out = dict(center=np.array(ap.center, dtype=float), ex=np.array(ap.xyz[0], dtype=float),
                   ey=np.array(ap.xyz[1], dtype=float), ez=np.array(ap.xyz[2], dtype=float),
                   lostNum=np.int64(ap.lostNum), r=float(ap.r), isBeamStop=stop, in_state_after=beam.state, lo_r=lo.r)
        out.update(beam_dict(beam_in, 'in_'))
        out.update(beam_dict(lo, 'out_'))
        st = lo.state
        print(name, {int(x): int((st == x).sum()) for x in np.unique(st)})
        save(name, **out)
    src = rs.GeometricSource(bl, 'src', nrays=10)
    np.random.seed(54)
    xy = np.random.rand(700, 2)
    np.random.seed(54)
    wave = ap.prepare_wave(src, 700)
    save('prepare_wave_round', xy=xy, r=float(ap.r), center=np.array(ap.center, dtype=float),
         ex=np.array(ap.xyz[0], dtype=float), ey=np.array(ap.xyz[1], dtype=float), ez=np.array(ap.xyz[2], dtype=float),
         sinAzimuth=bl.sinAzimuth, cosAzimuth=bl.cosAzimuth, src_center=np.array(src.center, dtype=float),
         x=wave.x, z=wave.z, area=wave.area, dS=wave.dS, xDiffr=wave.xDiffr, yDiffr=wave.yDiffr, zDiffr=wave.zDiffr,
         rDiffr=wave.rDiffr, a=wave.a, b=wave.b, c=wave.c)


def gen_figure_error():
    """OE.reflect with a figureError object (oes/base.py:681-686, :744-770; find_dz :825-829; reflect.py:767-776;
    figure_error.py:207-265): a Gaussian bump on a toroid, waviness on a flat mirror.  The bicubic spline of the height
    map (knots, coefficients) is stored as the reference's scipy RectBivariateSpline holds it."""
    import xrt.backends.raycing.figure_error as rfe
    bl = raycing.BeamLine(azimuth=0.1)
    Au = rm.Material('Au', rho=19.3)
    p = 2e4
    ctr = (p*np.sin(0.1), p*np.cos(0.1), 0)
    cases = (('toroid_bump', roe.ToroidMirror, dict(R=8e6, r=75.),
              lambda lx, ly: rfe.GaussianBump(bumpHeight=40., cX=0.5, cY=-30., sigmaX=1.5, sigmaY=60., limPhysX=lx,
                                              limPhysY=ly, gridStep=4.)),
             ('flat_waviness', roe.OE, dict(),
              lambda lx, ly: rfe.Waviness(amplitude=15., xWaveLength=3., yWaveLength=70., limPhysX=lx, limPhysY=ly,
                                          gridStep=4.)))
    for name, cls, kw, mkfe in cases:
        beam = _source_beam(bl, 2500, 401, dx=0.3, dz=0.1, dxprime=1.2e-4, dzprime=6e-5, energies=(7000, 11000))
        lx, ly = [-4, 4], [-350, 300]
        fe = mkfe(lx, ly)
        oe = cls(bl, name, center=ctr, pitch=3e-3, material=Au, limPhysX=lx, limPhysY=ly, figureError=fe, **kw)
        d = _capture_reflect(oe, beam)
        d.update(_oe_params(oe))
        spl = fe.local_z_spline
        ty, tx = spl.get_knots()
        st = d['gb_state']
        print(name, {int(x): int((st == x).sum()) for x in np.unique(st)}, 'knots', len(ty), len(tx), 'coeffs',
              spl.get_coeffs().shape, 'rms nm', float(fe.get_rms()))
        save('reflect_figerr_' + name, rho=19.3, R=float(getattr(oe, 'R', 0.)), r=float(getattr(oe, 'r', 0.)),
             fe_ty=ty, fe_tx=tx, fe_c=spl.get_coeffs(), fe_xShift=fe.xShift, fe_yShift=fe.yShift,
             fe_x1d=fe.x1d, fe_y1d=fe.y1d, fe_z=fe.generate_profile(), **d)


def gen_figure_error_param():
    """A figure error on a PARAMETRIC element (oes/base.py:790-797 local_r_distorted: the height map is added along z at
    the Cartesian image of (s, phi, r) and taken back to r; :744-770 local_n_distorted at that (x, y)): Gaussian bump on
    the cylindrical and on the ellipsoidal EllipticalMirrorParam."""
    import xrt.backends.raycing.figure_error as rfe
    bl = raycing.BeamLine()
    Au = rm.Material('Au', rho=19.3)
    for name, kw in (('ellipse_cyl_bump', dict(pitch=4e-3, p=30000., q=2000., isCylindrical=True)),
                     ('ellipsoid_bump', dict(pitch=5e-3, p=20000., q=5000., isCylindrical=False))):
        beam = _source_beam(bl, 2500, 411, dx=0.05, dz=0.03, dxprime=8e-6, dzprime=6e-6, energies=(8000, 10000))
        lx, ly = [-3, 3], [-90, 80]
        fe = rfe.GaussianBump(bumpHeight=25., cX=0.2, cY=-10., sigmaX=1.0, sigmaY=30., limPhysX=lx, limPhysY=ly,
                              gridStep=2.)
        oe = roe.EllipticalMirrorParam(bl, name, center=(0, kw['p'], 0), material=Au, limPhysX=lx, limPhysY=ly,
                                       figureError=fe, **kw)
        d = _capture_reflect(oe, beam)
        d.update(_oe_params(oe))
        for k in ('ellipseA', 'ellipseB', 'y0', 'z0', 'cosGamma', 'sinGamma', 'p', 'q'):
            d[k] = float(getattr(oe, k))
        d['isCylindrical'] = bool(oe.isCylindrical)
        spl = fe.local_z_spline
        ty, tx = spl.get_knots()
        st = d['gb_state']
        print(name, {int(x): int((st == x).sum()) for x in np.unique(st)}, 'knots', len(ty), len(tx), 'rms nm',
              float(fe.get_rms()))
        save('reflect_figerr_' + name, rho=19.3, fe_ty=ty, fe_tx=tx, fe_c=spl.get_coeffs(), fe_xShift=fe.xShift,
             fe_yShift=fe.yShift, fe_x1d=fe.x1d, fe_y1d=fe.y1d, fe_z=fe.generate_profile(), **d)


def gen_figure_error_dcm():
    """A figure error on a TWO-SURFACE element: DCM(figureError=...) -- find_dz adds the height map to local_z1 AND
    local_z2 (oes/base.py:825-829, each in its own crystal's local frame), _reflect_local turns the SURFACE normal of both
    crystals (reflect.py:767-776; with an asymmetric cut the normal of the planes stays)."""
    import xrt.backends.raycing.figure_error as rfe
    bl = raycing.BeamLine()
    si = rm.CrystalSi(hkl=(1, 1, 1))
    E0 = 9000.
    lx, ly = [-5, 5], [-60, 60]
    cases = (('waviness', None, lambda: rfe.Waviness(amplitude=8., xWaveLength=4., yWaveLength=25., limPhysX=lx,
                                                     limPhysY=ly, gridStep=2.)),
             # (with an asymmetric cut the reference itself fails: the normal of the planes is a tuple of scalars, the
             # turned surface normal one of arrays -> np.asarray raises at reflect.py:790)
             ('bump', None, lambda: rfe.GaussianBump(bumpHeight=20., cX=0.3, cY=5., sigmaX=2., sigmaY=15., limPhysX=lx,
                                                     limPhysY=ly, gridStep=2.)))
    for name, alpha, mkfe in cases:
        beam = _source_beam(bl, 3000, 431, dx=0.2, dz=0.02, dxprime=5e-5, dzprime=2e-5, energies=(8995, 9005))
        theta = si.get_Bragg_angle(E0) - si.get_dtheta(E0, alpha)
        fe = mkfe()
        dcm = roe.DCM(bl, 'dcm', center=(0, 2e4, 0), bragg=float(theta), material=si, material2=si, alpha=alpha,
                      cryst2perpTransl=10., cryst2finePitch=2e-6, cryst1roll=1e-4, figureError=fe,
                      limPhysX=lx, limPhysY=ly, limPhysX2=lx, limPhysY2=[-40, 80])
        gb, lo1, lo2 = dcm.double_reflect(beam)
        d = _dcm_params(dcm)
        d.update(beam_dict(beam, 'in_'))
        d.update(beam_dict(gb, 'gb_'))
        d.update(beam_dict(lo1, 'lo1_'))
        d.update(beam_dict(lo2, 'lo2_'))
        dcm0 = roe.DCM(bl, 'dcm0', center=(0, 2e4, 0), bragg=float(theta), material=si, material2=si, alpha=alpha,
                       cryst2perpTransl=10., cryst2finePitch=2e-6, cryst1roll=1e-4,
                       limPhysX=lx, limPhysY=ly, limPhysX2=lx, limPhysY2=[-40, 80])
        gb0, _, _ = dcm0.double_reflect(beam)
        both = (gb.state == 1) & (gb0.state == 1)
        print('dcm_figerr_' + name, {int(s): int((gb.state == s).sum()) for s in np.unique(gb.state)},
              'max |dc| vs perfect', float(np.abs(gb.c - gb0.c)[both].max()),
              'intensity ratio', float((gb.Jss + gb.Jpp)[both].sum() / (gb0.Jss + gb0.Jpp)[both].sum()),
              'rms nm', float(fe.get_rms()))
        save('dcm_figerr_' + name, d=si.d, tK=si.tK, E0=E0, fe_x1d=fe.x1d, fe_y1d=fe.y1d, fe_z=fe.generate_profile(),
             fe_c=fe.local_z_spline.get_coeffs(), fe_xShift=fe.xShift, fe_yShift=fe.yShift, **d)


def gen_figure_error_plate():
    """A figure error on a Plate (oes/refractive.py:12-234 derives from DCM): the same height map on the entrance and the
    exit surface (find_dz, base.py:825-829) and both refraction normals turned (reflect.py:767-776, toWhere = 1)."""
    import xrt.backends.raycing.figure_error as rfe
    bl = raycing.BeamLine(azimuth=0.1)
    ctr = (1e4*np.sin(0.1), 1e4*np.cos(0.1), 0)
    beam = _source_beam(bl, 2500, 441, dx=0.4, dz=0.3, dxprime=2e-5, dzprime=2e-5, energies=(7000, 11000))
    C = rm.Material('C', rho=3.52, kind='plate')
    lx, ly = [-0.9, 1.1], [-1.2, 1.0]
    fe = rfe.Waviness(amplitude=2000., xWaveLength=0.5, yWaveLength=0.7, limPhysX=lx, limPhysY=ly, gridStep=0.05)
    pl = roe.Plate(bl, 'W', center=ctr, pitch=np.radians(50.), t=0.3, wedgeAngle=2e-3, material=C, limPhysX=lx,
                   limPhysY=ly, figureError=fe)
    d = _capture_refract(pl, beam)
    pl0 = roe.Plate(bl, 'W0', center=ctr, pitch=np.radians(50.), t=0.3, wedgeAngle=2e-3, material=C, limPhysX=lx,
                    limPhysY=ly)
    gb0 = pl0.double_refract(beam)[0]
    both = (d['gb_state'] == 1) & (gb0.state == 1)
    st = d['gb_state']
    print('refract_plate_figerr', {int(x): int((st == x).sum()) for x in np.unique(st)}, 'max |dc| vs perfect',
          float(np.abs(d['gb_c'] - gb0.c)[both].max()), 'rms nm', float(fe.get_rms()))
    save('refract_plate_figerr', fe_x1d=fe.x1d, fe_y1d=fe.y1d, fe_z=fe.generate_profile(),
         fe_c=fe.local_z_spline.get_coeffs(), fe_xShift=fe.xShift, fe_yShift=fe.yShift, **d)


def gen_figure_error_lens():
    """A figure error on a parabolic lens stack (oes/refractive.py:237-549, multiple_refract :457-549): every lens of the
    CRL carries the same height map on both of its surfaces."""
    import xrt.backends.raycing.figure_error as rfe
    bl = raycing.BeamLine()
    Be = rm.Material('Be', rho=1.848, kind='lens')
    lx, ly = [-0.6, 0.6], [-0.6, 0.6]
    beam = _source_beam(bl, 2500, 451, dx=0.2, dz=0.2, dxprime=1e-5, dzprime=1e-5, energies=(8500, 9500))
    fe = rfe.Waviness(amplitude=1500., xWaveLength=0.35, yWaveLength=0.5, limPhysX=lx, limPhysY=ly, gridStep=0.03)
    kw = dict(center=(0, 10000, 0), pitch=np.pi/2, t=0.03, material=Be, focus=0.1, zmax=0.5, nCRL=3, limPhysX=lx,
              limPhysY=ly)
    ln = roe.DoubleParaboloidLens(bl, 'CRL', figureError=fe, **kw)
    d = _capture_refract(ln, beam, True)
    gb0 = roe.DoubleParaboloidLens(bl, 'CRL0', **kw).multiple_refract(beam)[0]
    st = d['gb_state']
    both = (st == 1) & (gb0.state == 1)
    print('refract_crl_figerr', {int(x): int((st == x).sum()) for x in np.unique(st)}, 'max |dc| vs perfect',
          float(np.abs(d['gb_c'] - gb0.c)[both].max()), 'rms nm', float(fe.get_rms()))
    save('refract_crl_figerr', focus=0.1, zmax=0.5, nCRL=np.int64(3), fe_x1d=fe.x1d, fe_y1d=fe.y1d,
         fe_z=fe.generate_profile(), fe_c=fe.local_z_spline.get_coeffs(), fe_xShift=fe.xShift, fe_yShift=fe.yShift, **d)


def gen_coated():
    """Coated (materials/multilayer.py:528-584): one reflective layer on a substrate, with surface and substrate
    roughness, kind 'mirror' -- amplitudes on a grid and a toroid mirror in a beam line."""
    Pt, Si = rm.Material('Pt', rho=21.45), rm.Material('Si', rho=2.33)
    kw = dict(coating=Pt, cThickness=350., surfaceRoughness=3., substrate=Si, substRoughness=5.)
    cm = rm.Coated(**kw)
    rng = np.random.RandomState(171)
    n = 2500
    theta = np.linspace(0.5e-3, 12e-3, n)
    E = rng.uniform(6000, 12000, n)
    bdn = -np.sin(theta)
    rs, rp = cm.get_amplitude(E.copy(), bdn.copy(), True)
    print('coated |rs|^2 range', float((abs(rs)**2).min()), float((abs(rs)**2).max()))
    save('coated_amplitudes', E=E, bdn=bdn, rs=rs, rp=rp, **{k: v for k, v in kw.items() if isinstance(v, float)})
    bl = raycing.BeamLine()
    beam = _source_beam(bl, 3000, 172, dx=0.5, dz=0.3, dxprime=2e-4, dzprime=1e-4, energies=(6000, 12000))
    oe = roe.ToroidMirror(bl, 'TM', center=(0, 25000, 0), pitch=3e-3, R=8e6, r=75., material=rm.Coated(**kw),
                          limPhysX=[-4, 4], limPhysY=[-500, 400])
    d = _capture_reflect(oe, beam)
    d.update(_oe_params(oe))
    st = d['gb_state']
    print('coated toroid', {int(x): int((st == x).sum()) for x in np.unique(st)}, 'mean R',
          float((d['gb_Jss'] + d['gb_Jpp'])[st == 1].mean() / (d['in_Jss'] + d['in_Jpp'])[st == 1].mean()))
    save('reflect_coated_toroid', R=float(oe.R), r=float(oe.r),
         **{k: v for k, v in kw.items() if isinstance(v, float)}, **d)


def gen_laue():
    """LauePlate and BentLaueCylinder (oes/laue.py:11-24, :26-175): Laue-geometry crystals with the planes' normal
    (n_B) and the surface normal returned together; Si(111) 'Laue reflected', with and without asymmetry angle."""
    bl = raycing.BeamLine()
    E0 = 9000.
    cases = (('laueplate', roe.LauePlate, dict(), None), ('laueplate_alpha', roe.LauePlate, dict(), np.radians(6.)),
             ('bentlaue_parab', roe.BentLaueCylinder, dict(R=2.5e4, crossSection='parabolic'), None),
             ('bentlaue_circ_alpha', roe.BentLaueCylinder, dict(R=2.5e4, crossSection='circular'), np.radians(-5.)))
    for i, (name, cls, kw, alpha) in enumerate(cases):
        si = rm.CrystalSi(hkl=(1, 1, 1), geom='Laue reflected', t=0.1)
        beam = _source_beam(bl, 1500, 230 + i, dx=0.2, dz=0.02, dxprime=5e-5, dzprime=2e-5, energies=(8998, 9002))
        thetaB = float(si.get_Bragg_angle(E0))
        pitch = np.pi/2 - thetaB + (alpha or 0.)
        oe = cls(bl, name, center=(0, 2e4, 0), pitch=pitch, material=si, alpha=alpha, limPhysX=[-20, 20],
                 limPhysY=[-20, 20], **kw)
        dd = _capture_reflect(oe, beam)
        dd.update(_oe_params(oe))
        st = dd['gb_state']
        good = st == 1
        print(name, {int(x): int((st == x).sum()) for x in np.unique(st)}, 'mean R',
              float((dd['gb_Jss'] + dd['gb_Jpp'])[good].mean() / (dd['in_Jss'] + dd['in_Jpp'])[good].mean()))
        save('reflect_laue_' + name, d=si.d, tK=si.tK, t=0.1, alpha=0. if alpha is None else alpha,
             R=float(kw.get('R', 0.)), **dd)


def gen_fzp():
    """NormalFZP (oes/gratings.py:10-137) with a material of kind 'FZP' (toWhere = 4, reflect.py:840-861): opaque
    zones absorb, transparent zones deflect by the local radial groove density; first and third order."""
    bl = raycing.BeamLine()
    for name, order, black in (('fzp_o1', 1, True), ('fzp_o3_white', 3, False)):
        mat = rm.Material('Au', rho=19.3, kind='FZP')
        beam = _source_beam(bl, 3000, 241, dx=0.02, dz=0.02, dxprime=2e-6, dzprime=2e-6, energies=(8990, 9010))
        oe = roe.NormalFZP(bl, name, center=(0, 5000, 0), pitch=np.pi/2, material=mat, f=100., E=9000., N=220,
                           isCentralZoneBlack=black, order=order)
        d = _capture_reflect(oe, beam)
        d.update(_oe_params(oe))
        st = d['gb_state']
        print(name, {int(x): int((st == x).sum()) for x in np.unique(st)}, 'rN', float(oe.rn[-1]))
        save('reflect_' + name, f=float(oe.f), E0=float(oe.E), N=int(oe.N), black=black, order=order, rn=oe.rn, **d)


def gen_multilayer_tran():
    """Multilayer(geom='transmitted') (materials/multilayer.py:388-516): transmitted amplitudes through the stack and
    its (thin) substrate -- on a grid and for a free-standing multilayer in a beam line (toWhere = 2: the rays pass
    straight, reflect.py:756, :862-878); cf. tests/raycing/test_multilayer_transmission.py."""
    Cr, C = rm.Material('Cr', rho=7.18), rm.Material('C', rho=2.26)
    kw = dict(tLayer=Cr, tThickness=15., bLayer=C, bThickness=25., nPairs=30, substrate=C, substThickness=100.,
              idThickness=1.5, geom='transmitted')
    ml = rm.Multilayer(**kw)
    rng = np.random.RandomState(181)
    n = 2500
    theta = np.linspace(np.radians(25.), np.radians(55.), n)
    E = rng.uniform(265., 275., n)
    bdn = -np.sin(theta)
    ts, tp = ml.get_amplitude(E.copy(), bdn.copy())
    print('ML transmission |ts|^2 range', float((abs(ts)**2).min()), float((abs(ts)**2).max()))
    save('multilayer_tran_amplitudes', E=E, bdn=bdn, ts=ts, tp=tp, dti=ml.dti, dbi=ml.dbi)
    bl = raycing.BeamLine()
    beam = _source_beam(bl, 2500, 182, dx=0.5, dz=0.01, dxprime=1e-4, dzprime=2e-2, energies=(268., 272.))
    oe = roe.OE(bl, 'mlt', center=(0, 100., 0), pitch=np.radians(35.), material=rm.Multilayer(**kw),
                limPhysX=[-5, 5], limPhysY=[-4, 4])
    dd = _capture_reflect(oe, beam)
    dd.update(_oe_params(oe))
    st = dd['gb_state']
    print('ML transmission plate', {int(x): int((st == x).sum()) for x in np.unique(st)}, 'mean T',
          float((dd['gb_Jss'] + dd['gb_Jpp'])[st == 1].mean() / (dd['in_Jss'] + dd['in_Jpp'])[st == 1].mean()))
    save('reflect_multilayer_tran', **dd)


def gen_hemispheric():
    """HemisphericScreen.expose (screens.py:422-545): a beam scattered into a wide cone, exposed on a sphere; some
    rays start outside the sphere and fly away from it (no intersection: lost)."""
    bl = raycing.BeamLine(azimuth=0.15)
    beam = _source_beam(bl, 2500, 61, dx=0.4, dz=0.3, dxprime=0.3, dzprime=0.2)
    far = np.random.RandomState(62).rand(2500) < 0.1
    beam.x[far] += 5000.  # outside the sphere
    scr = rsc.HemisphericScreen(bl, 'hemi', center=(30*np.sin(0.15), 30*np.cos(0.15), 0.5), R=400., phiOffset=-0.1,
                                thetaOffset=0.05)
    blo = scr.expose(beam)
    st = blo.state
    print('hemispheric', {int(x): int((st == x).sum()) for x in np.unique(st)})
    out = dict(center=np.array(scr.center, dtype=float), ex=np.array(scr.x, dtype=float), ey=np.array(scr.y, dtype=float),
               ez=np.array(scr.z, dtype=float), lostNum=np.int64(scr.lostNum), R=float(scr.R),
               phiOffset=float(scr.phiOffset), thetaOffset=float(scr.thetaOffset), sinAzimuth=bl.sinAzimuth,
               cosAzimuth=bl.cosAzimuth, out_phi=blo.phi)
    out.update(beam_dict(beam, 'in_'))
    out.update(beam_dict(blo, 'out_'))
    save('expose_hemispheric', **out)


def gen_blazed():
    """BlazedGrating (oes/gratings.py:316-535), constant line density: sawtooth surface with its own analytic
    find_intersection (first facet hit, shadowing respected) and facet normals; material kind 'mirror'."""
    bl = raycing.BeamLine()
    Au = rm.Material('Au', rho=19.3, kind='mirror')
    # (a variable line density cannot be generated: BlazedGrating(gratingDensity=[...]) recurses without end in the
    # reference as it stands -- reset() assigns self.rho0, whose setter calls reset(), gratings.py:377-380 / :416-420)
    for name, kw in (('blazed', dict(blaze=np.radians(1.2), rho=300.)),
                     ('blazed_anti', dict(blaze=np.radians(0.8), antiblaze=np.radians(30.), rho=600.))):
        beam = _source_beam(bl, 2500, 251, dx=0.3, dz=0.02, dxprime=1e-4, dzprime=3e-5, energies=(395., 405.))
        oe = roe.BlazedGrating(bl, name, center=(0, 15000., 0), pitch=np.radians(2.), material=Au, limPhysX=[-5, 5],
                               limPhysY=[-40, 45], **kw)
        d = _capture_reflect(oe, beam)
        d.update(_oe_params(oe))
        st = d['gb_state']
        print(name, {int(x): int((st == x).sum()) for x in np.unique(st)})
        save('reflect_' + name, rho=19.3, blaze=float(oe.blaze), antiblaze=float(oe.antiblaze), rho0=float(oe.rho0), **d)


def gen_roughness():
    """RandomRoughness height maps (figure_error.py:463-622) for seeded generators: RMS height, isotropic RMS slope and
    directional RMS slopes, one of them on top of a Gaussian bump (baseFE)."""
    import xrt.backends.raycing.figure_error as rfe
    lims = dict(limPhysX=[-4, 4], limPhysY=[-350, 300], gridStep=6.)
    out = {}
    for name, kw in (('height', dict(rms=2.5, rmsKind='height', corrLength=7., seed=12345)),
                     ('slope', dict(rms=0.8, rmsKind='slope', corrLength=12., seed=777)),
                     ('slope_dir', dict(rms=(0.5, 1.5), rmsKind='slope', corrLength=15., seed=4242))):
        base = rfe.GaussianBump(bumpHeight=10., sigmaX=2., sigmaY=80., **lims) if name == 'slope' else None
        fe = rfe.RandomRoughness(baseFE=base, **kw, **lims)
        out[name + '_z'] = fe.generate_profile()
        out[name + '_c'] = fe.local_z_spline.get_coeffs()
        print(name, 'rms nm', float(fe.get_rms()), 'rms slopes', [float(v) for v in fe.get_rms_slope()])
    save('figure_error_roughness', **out)


SECTIONS = {'figure_error_lens': gen_figure_error_lens, 'figure_error_plate': gen_figure_error_plate, 'figure_error_dcm': gen_figure_error_dcm, 'figure_error_param': gen_figure_error_param, 'roughness': gen_roughness, 'blazed': gen_blazed, 'hemispheric': gen_hemispheric, 'multilayer_tran': gen_multilayer_tran, 'fzp': gen_fzp, 'laue': gen_laue, 'coated': gen_coated, 'figure_error': gen_figure_error, 'round': gen_round, 'shine_wave': gen_shine_wave, 'multi': gen_multi, 'spliced': gen_spliced, 'field_source': gen_field_source, 'config1': gen_config1, 'mesh': gen_mesh, 'geometric': gen_geometric, 'refractive': gen_refractive, 'bm': gen_bm, 'gratings': gen_gratings, 'undulator_ext': gen_undulator_ext, 'thin': gen_thin_crystals, 'multilayer': gen_multilayer, 'hist': gen_hist, 'wave_oe': gen_wave_oe, 'diffract_oe': gen_diffract_oe, 'brent': gen_brent, 'diffract': gen_diffract, 'kb': gen_kb, 'shine': gen_shine, 'dcm': gen_dcm, 'amplitudes': gen_amplitudes, 'planes': gen_planes, 'reflect': gen_reflect, 'constants': gen_constants, 'kirchhoff': gen_kirchhoff, 'undulator': gen_undulator}

if __name__ == '__main__':
    todo = sys.argv[1:] or list(SECTIONS)
    for s in todo:
        SECTIONS[s]()
