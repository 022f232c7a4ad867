"""a13/a14: DCM.double_reflect with Si(111), symmetric and 5 deg asymmetric cut -- oracle bit-exact on the
reference's outputs (CPU), fused CUDA kernels vs golden (GPU)."""
import numpy as np
import pytest

from conftest import golden
from helpers import oracle_spec, oracle_beam, product_beamline, product_oe_kwargs, device_beam, compare_beams, \
    BEAM_FIELDS
from oracle import reflect as orf, materials as om

CASES = ['dcm_si111', 'dcm_si111_asym']
OUT_FIELDS = BEAM_FIELDS + ('theta',)
DCM_KEYS = ('bragg', 'cryst1roll', 'cryst2roll', 'cryst2pitch', 'cryst2finePitch', 'cryst2perpTransl',
            'cryst2longTransl')


@pytest.mark.parametrize('name', CASES)
def test_oracle_matches_reference_bit_exact(name):
    g = golden(name)
    si = om.CrystalDiamond((1, 1, 1), float(g['d']), 'Si')
    sp = oracle_spec(g, 'flat', si, alpha=float(g['alpha']) or None, material2=si,
                     surfPhysX2=list(g['surfPhysX2']), surfPhysY2=list(g['surfPhysY2']), surfOptX2=None,
                     surfOptY2=None, **{k: float(g[k]) for k in DCM_KEYS})
    gb, lo1, lo2 = orf.double_reflect(sp, oracle_beam(g, 'in_'))
    for pre, b in (('gb_', gb), ('lo1_', lo1), ('lo2_', lo2)):
        for f in OUT_FIELDS:
            if pre + f in g.files:
                assert np.array_equal(getattr(b, f), g[pre + f]), pre + f


@pytest.mark.gpu
@pytest.mark.parametrize('name', CASES)
def test_cuda_double_reflect_vs_golden(name, cuda_device):
    from xrt_b200 import oes, materials
    g = golden(name)
    si = materials.CrystalSi(hkl=(1, 1, 1), tK=float(g['tK']))
    assert si.d == float(g['d'])
    kw = product_oe_kwargs(g)
    kw.update({k: float(g[k]) for k in DCM_KEYS})
    dcm = oes.DCM(product_beamline(g), 'dcm', material=si, material2=si, alpha=float(g['alpha']) or None,
                  limPhysX2=list(g['surfPhysX2']), limPhysY2=list(g['surfPhysY2']), **kw)
    dcm.lostNum = int(g['lostNum'])
    gb, lo1, lo2 = dcm.double_reflect(device_beam(oracle_beam(g, 'in_')))
    bad = []
    for pre, b in (('gb_', gb), ('lo1_', lo1), ('lo2_', lo2)):
        ref = {f: g[pre + f] for f in OUT_FIELDS if pre + f in g.files}
        # the Bragg peak amplifies ulp-level differences of the transcendental functions like the Fresnel
        # case does: 1e-11 on the amplitude groups, 1e-12 on geometry
        bad += compare_beams(b.to_numpy(), ref, ('x', 'y', 'z', 'a', 'b', 'c', 'path', 'E', 'state', 'theta'),
                             1e-12, pre)
        bad += compare_beams(b.to_numpy(), ref, ('Jss', 'Jpp', 'Jsp', 'Es', 'Ep'), 1e-11, pre)
    assert not bad, bad


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_cuda_config3_at_1e6_rays_vs_oracle(cuda_device):
    """BASELINE.json configs[2] at 1e6 seeded rays (SURVEY 8d 'Config 3'): GeometricSource-like beam -> DCM Si(111)
    double_reflect, every ray against the numpy oracle -- the size where rare branches (rays at the crystal edges, deep
    in the Darwin tails, dead on entry) all occur; the goldens of the reference pin the oracle at 3000 rays."""
    from xrt_b200 import oes, materials, synthetic
    from xrt_b200.beamline import BeamLine
    n = 1000000
    g = synthetic.geometric_beam(4242, n, dx=0.2, dz=0.02, dxprime=5e-5, dzprime=2e-5, energies=(8995., 9005.),
                                 kill=0.01)
    si = materials.CrystalSi(hkl=(1, 1, 1))
    E0 = np.array([9000.])
    theta = float(si.get_Bragg_angle(E0)[0] - si.get_dtheta(E0)[0])
    bl = BeamLine()
    dcm = oes.DCM(bl, 'dcm', center=(0, 2e4, 0), bragg=theta, material=si, material2=si, cryst2perpTransl=10.,
                  limPhysX=[-0.5, 0.5], limPhysY=[-100, 100], limPhysX2=[-50, 50], limPhysY2=[-100, 200])
    gb, lo1, lo2 = dcm.double_reflect(device_beam(g))
    osi = om.CrystalDiamond((1, 1, 1), si.d, 'Si')
    sp = orf.OESpec(center=[0, 2e4, 0], surface='flat', surfPhysX=[-0.5, 0.5], surfPhysY=[-100, 100], material=osi,
                    material2=osi, bragg=theta, cryst1roll=0., cryst2roll=0., cryst2pitch=0., cryst2finePitch=0.,
                    cryst2perpTransl=10., cryst2longTransl=0., surfPhysX2=[-50, 50], surfPhysY2=[-100, 200],
                    surfOptX2=None, surfOptY2=None, lostNum=dcm.lostNum)
    ogb, olo1, olo2 = orf.double_reflect(sp, _obeam(g))
    st = ogb.state
    assert (st == 1).sum() > 0.3 * n and (st != 1).sum() > 1000  # narrow first crystal: many rays lost at its edges
    bad = []
    for pre, b, o in (('gb_', gb, ogb), ('lo1_', lo1, olo1), ('lo2_', lo2, olo2)):
        dv = b.to_numpy()
        bad += compare_beams(dv, o, ('x', 'y', 'z', 'a', 'b', 'c', 'path', 'E', 'state', 'theta'), 1e-12, pre)
        bad += compare_beams(dv, o, ('Jss', 'Jpp', 'Jsp', 'Es', 'Ep'), 1e-11, pre)
    assert not bad, bad


def _obeam(d):
    from oracle import beam as ob
    b = ob.Beam(len(d['x']), withAmplitudes='Es' in d)
    for k, v in d.items():
        setattr(b, k, np.array(v))
    return b


@pytest.mark.gpu
def test_cuda_asymmetric_crystal_met_from_both_sides_is_refused(cuda_device):
    """Documented deviation: the reference takes the sign of the asymmetric-cut deflection from the MEAN
    of n_B.k over the beam (reflect.py:575-576), the kernel per ray.  They agree whenever the whole beam meets the Bragg
    planes from one side -- every golden and config does; a beam that straddles the planes is detected (reduction
    slots 37 / 38) and refused, not silently deviated."""
    from xrt_b200 import oes, materials, synthetic
    from xrt_b200.beamline import BeamLine
    si = materials.CrystalSi(hkl=(1, 1, 1))
    theta = float(si.get_Bragg_angle(np.array([9000.]))[0])
    alpha = np.radians(30.)  # n_B tilted by 30 deg: grazing rays can meet the planes from either side
    n = 4000
    g = synthetic.geometric_beam(11, n, dx=0.05, dz=0.05, dxprime=1e-5, dzprime=1e-5, energies=(8999., 9001.), kill=0.)
    oe = oes.OE(BeamLine(), 'xtal', center=(0, 1e4, 0), pitch=theta + alpha, material=si, alpha=alpha,
                limPhysX=[-50, 50], limPhysY=[-500, 500])
    oe.reflect(device_beam(g))  # one side only: fine
    # second half of the beam climbs at an angle beta between the planes' pitch (theta) and the surface's pitch
    # (theta + alpha): it still meets the SURFACE from above (k.n_s = sin(beta - theta - alpha) < 0) but the Bragg
    # PLANES from below (k.n_B = sin(beta - theta) > 0), aiming at the crystal centre
    h = n // 2
    beta = theta + alpha / 2
    g['b'][h:] = np.cos(beta) * g['b'][h:] - np.sin(beta) * g['c'][h:]
    g['c'][h:] = np.sin(beta) * np.ones(n - h)
    norm = np.sqrt(g['a'][h:]**2 + g['b'][h:]**2 + g['c'][h:]**2)
    for k in 'abc':
        g[k][h:] /= norm
    g['z'][h:] -= 1e4 * np.tan(beta)
    with pytest.raises(NotImplementedError):
        oe.reflect(device_beam(g))
