"""f4: refractive elements -- Plate.double_refract and the parabolic lenses' multiple_refract
(oes/refractive.py:12-663; toWhere = 1 in oes/reflect.py:894-919; transmittivity material.py:482-486; absorption
and phase inside the material reflect.py:1048-1059).  Oracle pinned bit-exact on the reference's outputs (CPU),
CUDA kernels vs golden (GPU)."""
import numpy as np
import pytest

from conftest import golden
from helpers import oracle_spec, oracle_beam, product_beamline, product_oe_kwargs, device_beam, compare_beams, \
    BEAM_FIELDS
from oracle import reflect as orf, materials as om

CASES = [('refract_plate', 'Plate', 'C', 'plate'), ('refract_lens_single', 'ParaboloidFlatLens', 'Be', 'lens'),
         ('refract_crl_double', 'DoubleParaboloidLens', 'Be', 'lens'),
         ('refract_crl_cylinder', 'ParabolicCylinderFlatLens', 'Be', 'lens')]
OUT_FIELDS = BEAM_FIELDS + ('theta',)


def _oracle_run(name, cls, elem, kind):
    g = golden(name)
    mat = om.Material(elem, rho=float(g['rho']), kind=kind)
    kw = dict(bragg=0., cryst1roll=0., cryst2roll=0., cryst2pitch=float(g['wedgeAngle']), cryst2finePitch=0.,
              cryst2perpTransl=-float(g['t']), cryst2longTransl=0., material2=mat, t=float(g['t']),
              surfPhysX2=list(g['surfPhysX2']), surfPhysY2=list(g['surfPhysY2']), surfOptX2=None, surfOptY2=None)
    surface = 'flat'
    if cls != 'Plate':
        surface = 'paraboloid'
        kw.update(surfParams=dict(focus=float(g['focus']), zmax=float(g['zmax']), isCylinder='Cylinder' in cls),
                  nCRL=int(g['nCRL']), zmax=float(g['zmax']), isDouble=cls.startswith('Double'))
    sp = oracle_spec(g, surface, mat, **kw)
    beam = oracle_beam(g, 'in_')
    if cls == 'Plate':
        return g, orf.double_refract(sp, beam)
    return g, orf.multiple_refract(sp, beam)


@pytest.mark.parametrize('name,cls,elem,kind', CASES)
def test_oracle_matches_reference_bit_exact(name, cls, elem, kind):
    g, (gb, lo1, lo2) = _oracle_run(name, cls, elem, kind)
    for pre, b in (('gb_', gb), ('lo1_', lo1), ('lo2_', lo2)):
        for f in OUT_FIELDS:
            if pre + f in g.files:
                assert np.array_equal(getattr(b, f), g[pre + f]), pre + f
    st = g['gb_state']
    assert (st == 1).sum() > 1000 and (st < 0).sum() > 50


def _product_oe(g, cls, elem, kind):
    from xrt_b200 import oes, materials
    bl = product_beamline(g)
    kw = product_oe_kwargs(g)
    kw['material'] = materials.Material(elem, rho=float(g['rho']), kind=kind)
    kw['t'] = float(g['t'])
    if cls == 'Plate':
        oe = oes.Plate(bl, 'oe', wedgeAngle=float(g['wedgeAngle']), **kw)
    else:
        oe = getattr(oes, cls)(bl, 'oe', focus=float(g['focus']), zmax=float(g['zmax']), nCRL=int(g['nCRL']), **kw)
    oe.lostNum = int(g['lostNum'])
    assert list(oe.limPhysX2) == list(g['surfPhysX2']) and list(oe.limPhysY2) == list(g['surfPhysY2'])
    return oe


@pytest.mark.parametrize('name,cls,elem,kind', CASES)
def test_host_mirror_derives_the_reference_limits(name, cls, elem, kind):
    g = golden(name)
    oe = _product_oe(g, cls, elem, kind)
    assert oe.cryst2perpTransl == -float(g['t']) and oe.material2 is oe.material
    p2 = oe._make_params(0., 0., 0., is2ndXtal=True)
    assert [p2.physX[0], p2.physX[1]] == list(g['surfPhysX2'])


# Same operation order as the reference without FMA contraction; the transcendental functions (exp of the
# absorption, sincos of the phases, hypot of |n|) differ by <= 2 ulp: positions 1e-12 of the 1e4 mm scale,
# directions, coherency matrix and amplitudes 1e-12 of their group scale (Es, Ep carry exp(i k t) with k t ~ 5e11
# rad from the first surface: bit-identical t keeps that exact, as for mirrors).
RTOL = 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize('name,cls,elem,kind', CASES)
def test_cuda_refract_vs_reference_golden(name, cls, elem, kind, cuda_device):
    g = golden(name)
    oe = _product_oe(g, cls, elem, kind)
    beam = device_beam(oracle_beam(g, 'in_'))
    center0 = list(oe.center)
    gb, lo1, lo2 = oe.double_refract(beam) if cls == 'Plate' else oe.multiple_refract(beam)
    assert list(oe.center) == center0  # the CRL walk restores the centre
    bad = []
    for pre, b in (('gb_', gb), ('lo1_', lo1), ('lo2_', lo2)):
        ref = {f: g[pre + f] for f in OUT_FIELDS if pre + f in g.files}
        bad += compare_beams(b.to_numpy(), ref, OUT_FIELDS, RTOL, pre)
    assert not bad, bad


@pytest.mark.gpu
def test_cuda_crl_focuses(cuda_device):
    """Physics check independent of the reference: N double paraboloid Be lenses focus a parallel beam at
    f = focus / (2 N delta) * ... -- the rays' slopes are proportional to -x / f with f = R / (2 N delta), R = 2 focus."""
    from xrt_b200 import oes, materials
    from xrt_b200.beamline import BeamLine
    from xrt_b200.beams import Beam
    import torch
    n = 4000
    rng = np.random.RandomState(5)
    arr = dict(x=rng.uniform(-0.2, 0.2, n), y=np.zeros(n), z=rng.uniform(-0.2, 0.2, n), a=np.zeros(n), b=np.ones(n),
               c=np.zeros(n), path=np.zeros(n), E=np.full(n, 9000.), Jss=np.ones(n), Jpp=np.zeros(n),
               Jsp=np.zeros(n, dtype=complex), state=np.ones(n, dtype=np.int32))
    beam = Beam.from_numpy(arr)
    Be = materials.Material('Be', rho=1.848, kind='lens')
    N, focus = 4, 0.1
    lens = oes.DoubleParaboloidLens(BeamLine(), 'crl', center=(0, 1000, 0), pitch=np.pi/2, t=0.02, material=Be,
                                    focus=focus, zmax=0.4, nCRL=N, limPhysX=[-0.5, 0.5], limPhysY=[-0.5, 0.5])
    gb, _, _ = lens.multiple_refract(beam)
    out = gb.to_numpy()
    good = out['state'] == 1
    assert good.sum() == n
    delta = 1. - Be.get_refractive_index(np.array([9000.])).real[0]
    f = 2 * focus / (2 * N * delta)  # thin-lens focal length of N bi-concave parabolic lenses of apex radius 2 focus
    slope_x = out['a'][good] / out['b'][good]
    k = np.polyfit(arr['x'][good], slope_x, 1)[0]
    assert abs(k * f + 1) < 0.02  # thick-stack corrections are ~N * step / f
