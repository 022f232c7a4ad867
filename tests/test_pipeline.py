"""Config 5 chain at a size the oracle handles: Undulator (7th harmonic) -> DCM Si(111) -> vertical KB ->
horizontal KB -> Screen, every element on the GPU with device-resident beams, against the same chain through the
numpy oracle fed with the same random deviates."""
import numpy as np
import pytest

from helpers import compare_beams, device_beam, BEAM_FIELDS
from oracle import shine as osh, reflect as orf, planes as op, materials as om, undulator as ou

E1 = 1388.8614066632863
PITCH = 3.5e-3


def _geometry():
    si_d = 3.1354575567115175  # CrystalSi(111) at 297.15 K
    E0 = 7 * E1
    si = om.CrystalDiamond((1, 1, 1), si_d, 'Si')
    theta = float(si.get_Bragg_angle(np.array([E0]))[0] - si.get_dtheta(np.array([E0]))[0])
    g = dict(E0=E0, theta=theta, si_d=si_d, dcm_y=20000., vfm_y=30000., hfm_y=30600., scr_y=31800.)
    g['dcm_offset'] = 2 * 10. * np.cos(theta)  # cryst2perpTransl = 10 -> vertical beam offset 2 t cos(theta)
    return g


@pytest.mark.gpu
def test_cuda_config5_chain_vs_oracle(cuda_device):
    import torch
    from xrt_b200.beamline import BeamLine
    from xrt_b200.sources import Undulator
    from xrt_b200 import oes, materials
    from xrt_b200.planes import Screen
    geo = _geometry()
    n = 6000
    E0 = geo['E0']
    # ---------------- product objects
    bl = BeamLine()
    src = Undulator(bl, 'u', nrays=n, eE=3., eI=0.5, period=30., n=40, K=1.45, eSigmaX=48.65, eSigmaZ=6.197,
                    eEpsilonX=0.263, eEpsilonZ=0.008, xPrimeMax=0.03, zPrimeMax=0.03, xPrimeMaxAutoReduce=False,
                    zPrimeMaxAutoReduce=False, eMin=E0 - 1., eMax=E0 + 1., gNodes=64, gIntervals=2,
                    uniformRayDensity=True)
    src.reset()
    si = materials.CrystalSi(hkl=(1, 1, 1))
    assert si.d == geo['si_d']
    au = materials.Material('Au', rho=19.3)
    dcm = oes.DCM(bl, 'dcm', center=(0, geo['dcm_y'], 0), bragg=geo['theta'], material=si, material2=si,
                  cryst2perpTransl=10., limPhysX=[-20, 20], limPhysY=[-60, 60], limPhysX2=[-20, 20],
                  limPhysY2=[-60, 120])
    z1 = geo['dcm_offset']
    vfm = oes.EllipticalMirrorParam(bl, 'vfm', center=(0, geo['vfm_y'], z1), pitch=PITCH, p=geo['vfm_y'],
                                    q=geo['scr_y'] - geo['vfm_y'], isCylindrical=True, material=au,
                                    limPhysX=[-5, 5], limPhysY=[-150, 150])
    dyh = geo['hfm_y'] - geo['vfm_y']
    hfm = oes.EllipticalMirrorParam(bl, 'hfm', center=(0, geo['vfm_y'] + dyh * np.cos(2 * PITCH),
                                                      z1 + dyh * np.sin(2 * PITCH)),
                                    pitch=PITCH, positionRoll=np.pi / 2, p=geo['hfm_y'],
                                    q=geo['scr_y'] - geo['hfm_y'], isCylindrical=True, material=au,
                                    limPhysX=[-5, 5], limPhysY=[-150, 150], extraPitch=2 * PITCH)
    dys = geo['scr_y'] - geo['hfm_y']
    scr = Screen(bl, 'scr', center=(0, hfm.center[1] + dys * np.cos(2 * PITCH), hfm.center[2] + dys * np.sin(2 * PITCH)))
    # ---------------- oracle objects (same parameters)
    spec = osh.UndulatorSpec(L0=30., Np=40., Kx=0., Ky=1.45, eI=0.5, gamma=ou.gamma_from_eE(3.), phase=0., quadm=64,
                             gIntervals=2, E_min=src.E_min, E_max=src.E_max, Theta_min=src.Theta_min,
                             Theta_max=src.Theta_max, Psi_min=src.Psi_min, Psi_max=src.Psi_max, dx=src.dx, dz=src.dz,
                             dxprime=src.dxprime, dzprime=src.dzprime, uniformRayDensity=True)
    osi = om.CrystalDiamond((1, 1, 1), geo['si_d'], 'Si')
    oau = om.Material('Au', rho=19.3)
    odcm = orf.OESpec(center=list(dcm.center), surface='flat', surfPhysX=[-20, 20], surfPhysY=[-60, 60],
                      material=osi, material2=osi, bragg=geo['theta'], cryst1roll=0., cryst2roll=0., cryst2pitch=0.,
                      cryst2finePitch=0., cryst2perpTransl=10., cryst2longTransl=0., surfPhysX2=[-20, 20],
                      surfPhysY2=[-60, 120], surfOptX2=None, surfOptY2=None, lostNum=dcm.lostNum)

    def okb(o):
        return orf.OESpec(center=list(o.center), pitch=o.pitch, roll=o.roll + o.positionRoll, surface='ellipse_param',
                          isParametric=True, surfPhysX=[-5, 5], surfPhysY=[-150, 150], material=oau,
                          lostNum=o.lostNum, extraPitch=o.extraPitch,
                          surfParams=dict(ellipseA=o.ellipseA, ellipseB=o.ellipseB, y0=o.y0, z0=o.z0,
                                          cosGamma=o.cosGamma, sinGamma=o.sinGamma, isCylindrical=True))
    # ---------------- run both chains on the same deviates
    rec = []
    np.random.seed(2024)
    b0 = osh.shine(spec, n, record=rec)
    p = rec[0]
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    beam = src.shine(samples=dict(E=dev(p['E']), theta=dev(p['theta']), psi=dev(p['psi']), u=dev(np.zeros(n)),
                                  n1=dev(p['g1']), n2=dev(p['g2']), n3=dev(p['g3']), n4=dev(p['g4'])))
    o1, _, _ = orf.double_reflect(odcm, b0)
    d1, _, _ = dcm.double_reflect(beam)
    o2, _ = orf.reflect(okb(vfm), o1)
    d2, _ = vfm.reflect(d1)
    o3, _ = orf.reflect(okb(hfm), o2)
    d3, _ = hfm.reflect(d2)
    o4 = op.expose(o3, list(scr.center), [list(scr.x), list(scr.y), list(scr.z)], scr.lostNum)
    d4 = scr.expose(d3)
    out = d4.to_numpy()
    # the chain must transmit a decent part of the beam, otherwise the comparison is empty
    good = o4.state == 1
    assert good.sum() > 0.3 * n, good.sum()
    assert np.array_equal(out['state'], o4.state)
    for stage, dv, orc in (('dcm', d1, o1), ('vfm', d2, o2), ('hfm', d3, o3)):
        assert np.array_equal(dv.state.cpu().numpy(), orc.state), stage
    # geometry to 1e-12 of the 3e4 mm scale; amplitudes carry k*path ~ 1.5e12 rad and two parametric surfaces:
    # compare |E| and the coherency matrix (phase-free) tightly, the complex amplitudes modulo the path phase
    bad = compare_beams(out, o4, ('x', 'y', 'z', 'a', 'b', 'c', 'E', 'Jss', 'Jpp', 'Jsp'), 1e-10)
    assert not bad, bad
    assert np.abs(out['path'] - o4.path)[good].max() <= 2e-11
    from xrt_b200.physconsts import CHBAR
    comp = np.exp(-1e7j * o4.E / CHBAR * (out['path'] - o4.path))
    scale = np.abs(o4.Es).max()
    assert np.abs(out['Es'] * comp - o4.Es).max() <= 2e-3 * scale
    assert np.abs(np.abs(out['Es']) - np.abs(o4.Es)).max() <= 1e-10 * scale
