"""f4: figure errors -- OE(figureError=...) (oes/base.py:681-686, :744-770, find_dz :825-829, reflect.py:767-776;
figure_error.py:207-265): a height map (bicubic spline) added to the surface in the intersection search and its slopes
applied to the surface normal as d_pitch / d_roll rotations.  Oracle bit-exact on the reference's outputs (CPU); the
fused kernel with the spline evaluated on the device (GPU)."""
import numpy as np
import pytest

from conftest import golden
from helpers import (oracle_spec, oracle_beam, product_beamline, product_oe_kwargs, device_beam, compare_beams,
                     BEAM_FIELDS)
from oracle import materials as om, reflect as orf, surfaces as osf

CASES = [('toroid_bump', 'toroid'), ('flat_waviness', 'flat')]
OUT_FIELDS = BEAM_FIELDS + ('theta',)


@pytest.mark.parametrize('name,surface', CASES)
def test_oracle_figure_error_bit_exact(name, surface):
    g = golden('reflect_figerr_' + name)
    fe = osf.FigureErrorSpline(g['fe_x1d'], g['fe_y1d'], g['fe_z'], float(g['fe_xShift']), float(g['fe_yShift']))
    ty, tx = fe.local_z_spline.get_knots()
    assert np.array_equal(ty, g['fe_ty']) and np.array_equal(tx, g['fe_tx'])
    assert np.array_equal(fe.local_z_spline.get_coeffs(), g['fe_c'])
    kw = dict(surfParams=dict(R=float(g['R']), r=float(g['r']))) if surface == 'toroid' else {}
    sp = oracle_spec(g, surface, om.Material('Au', rho=float(g['rho'])), figureError=fe, **kw)
    info = {}
    gb, lb = orf.reflect(sp, oracle_beam(g, 'in_'), info=info)
    for pre, b in (('gb_', gb), ('lb_', lb)):
        for f in OUT_FIELDS:
            if pre + f in g.files:
                assert np.array_equal(getattr(b, f), g[pre + f]), pre + f
    assert np.array_equal(info['footprint'], g['footprint'])


PARAM_CASES = ['ellipse_cyl_bump', 'ellipsoid_bump']
ELLIPSE_KEYS = ('ellipseA', 'ellipseB', 'y0', 'z0', 'cosGamma', 'sinGamma')


@pytest.mark.parametrize('name', PARAM_CASES)
def test_oracle_figure_error_on_parametric_surface_bit_exact(name):
    """local_r_distorted / local_n_distorted of a parametric element (oes/base.py:790-797, :765-768): the height map is
    added along z at the Cartesian image of (s, phi, r) and taken back to r; the slopes are read at that (x, y)."""
    g = golden('reflect_figerr_' + name)
    fe = osf.FigureErrorSpline(g['fe_x1d'], g['fe_y1d'], g['fe_z'], float(g['fe_xShift']), float(g['fe_yShift']))
    assert np.array_equal(fe.local_z_spline.get_coeffs(), g['fe_c'])
    kw = dict(isParametric=True,
              surfParams=dict(isCylindrical=bool(g['isCylindrical']), **{k: float(g[k]) for k in ELLIPSE_KEYS}))
    mat = om.Material('Au', rho=float(g['rho']))
    gb, lb = orf.reflect(oracle_spec(g, 'ellipse_param', mat, figureError=fe, **kw), oracle_beam(g, 'in_'))
    for pre, b in (('gb_', gb), ('lb_', lb)):
        for f in OUT_FIELDS + ('s', 'phi', 'r'):
            if pre + f in g.files:
                assert np.array_equal(getattr(b, f), g[pre + f]), pre + f
    # ... and the bump matters: without it the directions differ
    gb0, _ = orf.reflect(oracle_spec(g, 'ellipse_param', mat, **kw), oracle_beam(g, 'in_'))
    good = (gb.state == 1) & (gb0.state == 1)
    assert 1e-10 < np.abs(gb.c[good] - gb0.c[good]).max() < 1e-5


def _product_oe(name, surface, g, fe):
    from xrt_b200 import oes, materials
    kw = product_oe_kwargs(g)
    kw['material'] = materials.Material('Au', rho=float(g['rho']))
    if surface == 'toroid':
        oe = oes.ToroidMirror(product_beamline(g), name, R=float(g['R']), r=float(g['r']), figureError=fe, **kw)
    else:
        oe = oes.OE(product_beamline(g), name, figureError=fe, **kw)
    oe.lostNum = int(g['lostNum'])
    return oe


def test_product_figure_error_profiles_and_splines_equal_reference():
    """Host side (no GPU): the product's GaussianBump / Waviness generate the reference's height maps on the
    reference's grids, and their splines have the reference's knots and coefficients bit for bit."""
    from xrt_b200 import figure_error as pfe
    g = golden('reflect_figerr_toroid_bump')
    fe = pfe.GaussianBump(bumpHeight=40., cX=0.5, cY=-30., sigmaX=1.5, sigmaY=60., limPhysX=[-4, 4],
                          limPhysY=[-350, 300], gridStep=4.)
    assert np.array_equal(fe.x1d, g['fe_x1d']) and np.array_equal(fe.y1d, g['fe_y1d'])
    assert np.array_equal(fe.generate_profile(), g['fe_z'])
    assert np.array_equal(fe.local_z_spline.get_coeffs(), g['fe_c'])
    g = golden('reflect_figerr_flat_waviness')
    fe = pfe.Waviness(amplitude=15., xWaveLength=3., yWaveLength=70., limPhysX=[-4, 4], limPhysY=[-350, 300],
                      gridStep=4.)
    assert np.array_equal(fe.generate_profile(), g['fe_z'])
    ty, tx = fe.local_z_spline.get_knots()
    assert np.array_equal(ty, g['fe_ty']) and np.array_equal(tx, g['fe_tx'])


@pytest.mark.gpu
@pytest.mark.parametrize('name,surface', CASES)
def test_cuda_figure_error_vs_reference_golden(name, surface, cuda_device):
    """The spline is evaluated on the device in FITPACK's operation order; what remains are the ulp-level differences of
    the de Boor recurrences, which move the solver's stopping point like any transcendental (as for parametric
    surfaces): |dt| <= 4e-12 mm, fields compared after removing exp(i k dt); everything else 1e-12."""
    from xrt_b200 import figure_error as pfe
    from xrt_b200.physconsts import CHBAR
    g = golden('reflect_figerr_' + name)
    fe = pfe.FigureErrorMap(g['fe_x1d'], g['fe_y1d'], g['fe_z'])
    oe = _product_oe(name, surface, g, fe)
    gb, lb = oe.reflect(device_beam(oracle_beam(g, 'in_')), withFootprint=True)
    bad = []
    good = g['lb_state'] == 1
    for pre, b in (('gb_', gb), ('lb_', lb)):
        ref = {f: g[pre + f] for f in OUT_FIELDS if pre + f in g.files}
        dev = b.to_numpy()
        dpath = dev['path'] - ref['path']
        assert np.abs(dpath[good]).max() <= 4e-12
        comp = np.exp(-1e7j * ref['E'] / CHBAR * dpath)
        dev = dict(dev, Es=dev['Es'] * comp, Ep=dev['Ep'] * comp)
        phase_ulp = np.spacing((1e7 * ref['E'] / CHBAR * ref['path']).max())
        bad += compare_beams(dev, ref, ('Es', 'Ep'), 2 * phase_ulp, pre)
        bad += compare_beams(dev, ref, [f for f in OUT_FIELDS if f not in ('Es', 'Ep')], 1e-12, pre)
    assert not bad, bad
    assert np.array_equal(gb.to_numpy()['state'], g['gb_state'])
    np.testing.assert_allclose(oe.footprint[0], g['footprint'], rtol=1e-12, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize('name', PARAM_CASES)
def test_cuda_figure_error_on_parametric_surface_vs_reference_golden(name, cuda_device):
    """EllipticalMirrorParam(figureError=...): the figure-error plug-in variant of the elliptic surface (kind 2) --
    local_r_distorted in find_dz, the slopes at the Cartesian image of (s, phi).  Tolerances as for the parametric
    surface without a figure error (tests/test_reflect.py) and for figure errors on z(x, y) surfaces (above)."""
    from xrt_b200 import figure_error as pfe, oes, materials
    from xrt_b200.physconsts import CHBAR
    g = golden('reflect_figerr_' + name)
    fe = pfe.FigureErrorMap(g['fe_x1d'], g['fe_y1d'], g['fe_z'])
    kw = product_oe_kwargs(g)
    oe = oes.EllipticalMirrorParam(product_beamline(g), name, p=float(g['p']), q=float(g['q']),
                                   isCylindrical=bool(g['isCylindrical']),
                                   material=materials.Material('Au', rho=float(g['rho'])), figureError=fe, **kw)
    for k in ELLIPSE_KEYS:
        assert abs(getattr(oe, k) - float(g[k])) <= 1e-13 * max(1., abs(float(g[k]))), k
        setattr(oe, k, float(g[k]))
    oe.lostNum = int(g['lostNum'])
    assert oe._plugin().figure_error and oe._plugin().kind == 2
    gb, lb = oe.reflect(device_beam(oracle_beam(g, 'in_')), withFootprint=True)
    assert np.array_equal(gb.to_numpy()['state'], g['gb_state'])
    bad = []
    good = g['lb_state'] == 1
    for pre, b in (('gb_', gb), ('lb_', lb)):
        ref = {f: g[pre + f] for f in OUT_FIELDS if pre + f in g.files}
        dev = b.to_numpy()
        dpath = dev['path'] - ref['path']
        assert np.abs(dpath[good]).max() <= 4e-12
        comp = np.exp(-1e7j * ref['E'] / CHBAR * dpath)
        dev = dict(dev, Es=dev['Es'] * comp, Ep=dev['Ep'] * comp)
        phase_ulp = np.spacing((1e7 * ref['E'] / CHBAR * ref['path']).max())
        bad += compare_beams(dev, ref, ('Es', 'Ep'), 2 * phase_ulp, pre)
        bad += compare_beams(dev, ref, [f for f in OUT_FIELDS if f not in ('Es', 'Ep')], 1e-12, pre)
    assert not bad, bad
    np.testing.assert_allclose(oe.footprint[0], g['footprint'], rtol=1e-12, atol=1e-12)
    # the bump is seen: the same element without it sends the rays elsewhere
    oe0 = oes.EllipticalMirrorParam(product_beamline(g), name + '0', p=float(g['p']), q=float(g['q']),
                                    isCylindrical=bool(g['isCylindrical']),
                                    material=materials.Material('Au', rho=float(g['rho'])), **kw)
    for k in ELLIPSE_KEYS:
        setattr(oe0, k, float(g[k]))
    oe0.lostNum = int(g['lostNum'])
    c0 = oe0.reflect(device_beam(oracle_beam(g, 'in_')))[0].to_numpy()
    both = (c0['state'] == 1) & (g['gb_state'] == 1)
    assert 1e-10 < np.abs(c0['c'][both] - g['gb_c'][both]).max() < 1e-5


@pytest.mark.gpu
def test_cuda_figure_error_changes_the_beam_and_library_refuses_it(cuda_device):
    """Sanity: the bump really deflects rays (the slopes are ~1e-7 rad), and the plain library object refuses a
    figure-error parameter block instead of ignoring it."""
    import ctypes
    from xrt_b200 import figure_error as pfe, _lib
    g = golden('reflect_figerr_toroid_bump')
    fe = pfe.FigureErrorMap(g['fe_x1d'], g['fe_y1d'], g['fe_z'])
    beam = device_beam(oracle_beam(g, 'in_'))
    with_fe = _product_oe('a', 'toroid', g, fe).reflect(beam)[0].to_numpy()
    plain_oe = _product_oe('b', 'toroid', g, None)
    plain = plain_oe.reflect(beam)[0].to_numpy()
    good = (with_fe['state'] == 1) & (plain['state'] == 1)
    d = np.abs(with_fe['c'][good] - plain['c'][good]).max()
    assert 1e-9 < d < 1e-5, d
    params = plain_oe._make_params(plain_oe.pitch, 0., 0.)
    fe.fill_params(params, beam.device)
    out = beam.like()
    cin, cout = beam.c_struct(), out.c_struct()
    red = plain_oe._red[str(beam.device)]
    rc = _lib.load().xrt_reflect_f64(len(beam), ctypes.byref(params), None, None, 0, ctypes.byref(cin), None,
                                     ctypes.byref(cout), None, red.data_ptr(), 7, 0, None)
    assert rc != 0 and b'figure' in _lib.load().xrt_b200_last_error()


@pytest.mark.parametrize('name,kw', [('height', dict(rms=2.5, rmsKind='height', corrLength=7., seed=12345)),
                                     ('slope', dict(rms=0.8, rmsKind='slope', corrLength=12., seed=777)),
                                     ('slope_dir', dict(rms=(0.5, 1.5), rmsKind='slope', corrLength=15., seed=4242))])
def test_product_random_roughness_equals_reference(name, kw):
    """Host side (no GPU): seeded RandomRoughness maps (RMS height, isotropic and directional RMS slope, one on top of a
    Gaussian bump) and their spline coefficients equal the reference's."""
    from xrt_b200 import figure_error as pfe
    g = golden('figure_error_roughness')
    lims = dict(limPhysX=[-4, 4], limPhysY=[-350, 300], gridStep=6.)
    base = pfe.GaussianBump(bumpHeight=10., sigmaX=2., sigmaY=80., **lims) if name == 'slope' else None
    fe = pfe.RandomRoughness(baseFE=base, **kw, **lims)
    z = fe.generate_profile()
    scale = np.abs(g[name + '_z']).max()
    assert np.abs(z - g[name + '_z']).max() <= 1e-13 * scale
    assert np.abs(fe.local_z_spline.get_coeffs() - g[name + '_c']).max() <= 1e-12 * scale
    if name == 'height':
        assert abs(fe.get_rms() - 2.5) < 1e-12


# ---------------------------------------------------------------- two-surface element: DCM(figureError=...)
DCM_CASES = ['waviness', 'bump']
DCM_KEYS = ('bragg', 'cryst1roll', 'cryst2roll', 'cryst2pitch', 'cryst2finePitch', 'cryst2perpTransl',
            'cryst2longTransl')


@pytest.mark.parametrize('name', DCM_CASES)
def test_oracle_figure_error_on_dcm_bit_exact(name):
    """DCM.double_reflect with a figure error (oes/dcm.py:247-354): find_dz adds the height map to local_z1 and to
    local_z2, each in its own crystal's frame (base.py:825-829); both crystals' normals are turned (reflect.py:767-776).
    The reference supports it for symmetric cuts only -- with alpha its own np.asarray(oeNormal) fails (scalars of the
    planes' normal beside the arrays of the turned surface normal, reflect.py:790)."""
    g = golden('dcm_figerr_' + name)
    fe = osf.FigureErrorSpline(g['fe_x1d'], g['fe_y1d'], g['fe_z'], float(g['fe_xShift']), float(g['fe_yShift']))
    assert np.array_equal(fe.local_z_spline.get_coeffs(), g['fe_c'])
    si = om.CrystalDiamond((1, 1, 1), float(g['d']), 'Si')
    kw = dict(material2=si, surfPhysX2=list(g['surfPhysX2']), surfPhysY2=list(g['surfPhysY2']), surfOptX2=None,
              surfOptY2=None, **{k: float(g[k]) for k in DCM_KEYS})
    gb, lo1, lo2 = orf.double_reflect(oracle_spec(g, 'flat', si, figureError=fe, **kw), oracle_beam(g, 'in_'))
    for pre, b in (('gb_', gb), ('lo1_', lo1), ('lo2_', lo2)):
        for f in OUT_FIELDS:
            if pre + f in g.files:
                assert np.array_equal(getattr(b, f), g[pre + f]), pre + f
    gb0, _, _ = orf.double_reflect(oracle_spec(g, 'flat', si, **kw), oracle_beam(g, 'in_'))  # ... and it matters
    good = (gb.state == 1) & (gb0.state == 1)
    assert good.sum() > 2000 and 1e-7 < np.abs(gb.c[good] - gb0.c[good]).max() < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize('name', DCM_CASES)
def test_cuda_figure_error_on_dcm_vs_reference_golden(name, cuda_device):
    """DCM(figureError=...) through the figure-error plug-in variant of the flat surface, both crystals.  Tolerances:
    geometry 1e-12 (|dt| <= 4e-12 mm per crystal as for the single-surface figure errors), amplitudes 1e-11 of the
    group scale after removing exp(i k dt) (the Bragg peak amplifies ulp-level differences, tests/test_dcm.py)."""
    from xrt_b200 import figure_error as pfe, oes, materials
    from xrt_b200.physconsts import CHBAR
    g = golden('dcm_figerr_' + name)
    fe = pfe.FigureErrorMap(g['fe_x1d'], g['fe_y1d'], g['fe_z'])
    si = materials.CrystalSi(hkl=(1, 1, 1), tK=float(g['tK']))
    kw = product_oe_kwargs(g)
    kw.update({k: float(g[k]) for k in DCM_KEYS})
    dcm = oes.DCM(product_beamline(g), 'dcm', material=si, material2=si, figureError=fe,
                  limPhysX2=list(g['surfPhysX2']), limPhysY2=list(g['surfPhysY2']), **kw)
    dcm.lostNum = int(g['lostNum'])
    assert dcm._plugin().figure_error and dcm._plugin().kind == 0
    gb, lo1, lo2 = dcm.double_reflect(device_beam(oracle_beam(g, 'in_')))
    assert np.array_equal(gb.to_numpy()['state'], g['gb_state'])
    bad = []
    for pre, b in (('gb_', gb), ('lo1_', lo1), ('lo2_', lo2)):
        ref = {f: g[pre + f] for f in OUT_FIELDS if pre + f in g.files}
        dev = b.to_numpy()
        good = ref['state'] == 1
        dpath = np.where(good, dev['path'] - ref['path'], 0.)
        assert np.abs(dpath).max() <= 8e-12, (pre, float(np.abs(dpath).max()))
        comp = np.exp(-1e7j * ref['E'] / CHBAR * dpath)
        dev = dict(dev, Es=dev['Es'] * comp, Ep=dev['Ep'] * comp)
        bad += compare_beams(dev, ref, ('x', 'y', 'z', 'a', 'b', 'c', 'path', 'E', 'state', 'theta'), 1e-12, pre)
        bad += compare_beams(dev, ref, ('Jss', 'Jpp', 'Jsp', 'Es', 'Ep'), 1e-11, pre)
    assert not bad, bad
    # an asymmetric cut with a figure error has no reference behaviour to match (above): refused
    dcm.alpha = np.radians(5.)
    with pytest.raises(NotImplementedError):
        dcm._plugin()


# ---------------------------------------------------------------- Plate(figureError=...): both refracting surfaces
def _plate_oracle(g, fe):
    mat = om.Material('C', rho=float(g['rho']), kind='plate')
    kw = dict(bragg=0., cryst1roll=0., cryst2roll=0., cryst2pitch=float(g['wedgeAngle']), cryst2finePitch=0.,
              cryst2perpTransl=-float(g['t']), cryst2longTransl=0., material2=mat, t=float(g['t']),
              surfPhysX2=list(g['surfPhysX2']), surfPhysY2=list(g['surfPhysY2']), surfOptX2=None, surfOptY2=None)
    if fe is not None:
        kw['figureError'] = fe
    return orf.double_refract(oracle_spec(g, 'flat', mat, **kw), oracle_beam(g, 'in_'))


def test_oracle_figure_error_on_plate_bit_exact():
    """Plate.double_refract with a figure error (oes/refractive.py:170-234): the height map displaces the entrance and
    the exit surface, Snell's law uses the turned normals (reflect.py:767-776, :894-919)."""
    g = golden('refract_plate_figerr')
    fe = osf.FigureErrorSpline(g['fe_x1d'], g['fe_y1d'], g['fe_z'], float(g['fe_xShift']), float(g['fe_yShift']))
    assert np.array_equal(fe.local_z_spline.get_coeffs(), g['fe_c'])
    gb, lo1, lo2 = _plate_oracle(g, fe)
    for pre, b in (('gb_', gb), ('lo1_', lo1), ('lo2_', lo2)):
        for f in OUT_FIELDS:
            if pre + f in g.files:
                assert np.array_equal(getattr(b, f), g[pre + f]), pre + f
    gb0 = _plate_oracle(g, None)[0]
    good = (gb.state == 1) & (gb0.state == 1)
    assert good.sum() > 2000 and 1e-8 < np.abs(gb.c[good] - gb0.c[good]).max() < 1e-5


@pytest.mark.gpu
def test_cuda_figure_error_on_plate_vs_reference_golden(cuda_device):
    """Plate(figureError=...) through the figure-error plug-in variant of the flat surface, both surfaces; tolerances of
    the single-surface figure-error tests (|dt| <= 4e-12 mm per surface, exp(i k dt) removed, the rest 1e-12)."""
    from xrt_b200 import figure_error as pfe, oes, materials
    from xrt_b200.physconsts import CHBAR
    g = golden('refract_plate_figerr')
    fe = pfe.FigureErrorMap(g['fe_x1d'], g['fe_y1d'], g['fe_z'])
    kw = product_oe_kwargs(g)
    kw['material'] = materials.Material('C', rho=float(g['rho']), kind='plate')
    oe = oes.Plate(product_beamline(g), 'W', wedgeAngle=float(g['wedgeAngle']), t=float(g['t']), figureError=fe, **kw)
    oe.lostNum = int(g['lostNum'])
    assert oe._plugin().figure_error and oe._plugin().kind == 0
    gb, lo1, lo2 = oe.double_refract(device_beam(oracle_beam(g, 'in_')))
    assert np.array_equal(gb.to_numpy()['state'], g['gb_state'])
    bad = []
    for pre, b in (('gb_', gb), ('lo1_', lo1), ('lo2_', lo2)):
        ref = {f: g[pre + f] for f in OUT_FIELDS if pre + f in g.files}
        dev = b.to_numpy()
        good = ref['state'] == 1
        dpath = np.where(good, dev['path'] - ref['path'], 0.)
        assert np.abs(dpath).max() <= 8e-12, (pre, float(np.abs(dpath).max()))
        comp = np.exp(-1e7j * ref['E'] / CHBAR * dpath)
        dev = dict(dev, Es=dev['Es'] * comp, Ep=dev['Ep'] * comp)
        phase_ulp = np.spacing((1e7 * ref['E'] / CHBAR * ref['path']).max())
        bad += compare_beams(dev, ref, ('Es', 'Ep'), 2 * phase_ulp, pre)
        bad += compare_beams(dev, ref, [f for f in OUT_FIELDS if f not in ('Es', 'Ep')], 1e-12, pre)
    assert not bad, bad


# ---------------------------------------------------------------- DoubleParaboloidLens(figureError=...), nCRL = 3
def _crl_oracle(g, fe):
    mat = om.Material('Be', rho=float(g['rho']), kind='lens')
    kw = dict(bragg=0., cryst1roll=0., cryst2roll=0., cryst2pitch=float(g['wedgeAngle']), cryst2finePitch=0.,
              cryst2perpTransl=-float(g['t']), cryst2longTransl=0., material2=mat, t=float(g['t']),
              surfPhysX2=list(g['surfPhysX2']), surfPhysY2=list(g['surfPhysY2']), surfOptX2=None, surfOptY2=None,
              surfParams=dict(focus=float(g['focus']), zmax=float(g['zmax']), isCylinder=False), nCRL=int(g['nCRL']),
              zmax=float(g['zmax']), isDouble=True)
    if fe is not None:
        kw['figureError'] = fe
    return orf.multiple_refract(oracle_spec(g, 'paraboloid', mat, **kw), oracle_beam(g, 'in_'))


def test_oracle_figure_error_on_lens_stack_bit_exact():
    """multiple_refract with a figure error (oes/refractive.py:457-549): the same height map on both surfaces of every
    lens of the stack."""
    g = golden('refract_crl_figerr')
    fe = osf.FigureErrorSpline(g['fe_x1d'], g['fe_y1d'], g['fe_z'], float(g['fe_xShift']), float(g['fe_yShift']))
    assert np.array_equal(fe.local_z_spline.get_coeffs(), g['fe_c'])
    gb, lo1, lo2 = _crl_oracle(g, fe)
    for pre, b in (('gb_', gb), ('lo1_', lo1), ('lo2_', lo2)):
        for f in OUT_FIELDS:
            if pre + f in g.files:
                assert np.array_equal(getattr(b, f), g[pre + f]), pre + f
    gb0 = _crl_oracle(g, None)[0]
    good = (gb.state == 1) & (gb0.state == 1)
    assert good.sum() > 2000 and 1e-7 < np.abs(gb.c[good] - gb0.c[good]).max() < 1e-4


@pytest.mark.gpu
def test_cuda_figure_error_on_lens_stack_vs_reference_golden(cuda_device):
    """DoubleParaboloidLens(figureError=..., nCRL=3).multiple_refract through the figure-error plug-in variant of the
    parabolic surface (kind 3): six refractions, each with the height map in the intersection search and the turned
    normal.  |dt| <= 4e-12 mm per surface, exp(i k dt) removed, the rest 1e-12 of the group scale."""
    from xrt_b200 import figure_error as pfe, oes, materials
    from xrt_b200.physconsts import CHBAR
    g = golden('refract_crl_figerr')
    fe = pfe.FigureErrorMap(g['fe_x1d'], g['fe_y1d'], g['fe_z'])
    kw = product_oe_kwargs(g)
    kw['material'] = materials.Material('Be', rho=float(g['rho']), kind='lens')
    oe = oes.DoubleParaboloidLens(product_beamline(g), 'CRL', t=float(g['t']), focus=float(g['focus']),
                                  zmax=float(g['zmax']), nCRL=int(g['nCRL']), figureError=fe, **kw)
    oe.lostNum = int(g['lostNum'])
    assert oe._plugin().figure_error and oe._plugin().kind == 3
    gb, lo1, lo2 = oe.multiple_refract(device_beam(oracle_beam(g, 'in_')))
    assert np.array_equal(gb.to_numpy()['state'], g['gb_state'])
    bad = []
    for pre, b, nsurf in (('gb_', gb, 6), ('lo1_', lo1, 1), ('lo2_', lo2, 2)):
        ref = {f: g[pre + f] for f in OUT_FIELDS if pre + f in g.files}
        dev = b.to_numpy()
        good = ref['state'] == 1
        dpath = np.where(good, dev['path'] - ref['path'], 0.)
        assert np.abs(dpath).max() <= 4e-12 * nsurf, (pre, float(np.abs(dpath).max()))
        comp = np.exp(-1e7j * ref['E'] / CHBAR * dpath)
        dev = dict(dev, Es=dev['Es'] * comp, Ep=dev['Ep'] * comp)
        phase_ulp = np.spacing((1e7 * ref['E'] / CHBAR * ref['path']).max())
        bad += compare_beams(dev, ref, ('Es', 'Ep'), 2 * nsurf * phase_ulp, pre)
        bad += compare_beams(dev, ref, [f for f in OUT_FIELDS if f not in ('Es', 'Ep')], 1e-12, pre)
    assert not bad, bad


def test_host_figure_error_plugin_selection_and_refusals():
    """Host side (no GPU): which plug-in object serves an element with a figure error, and what is refused -- an
    asymmetric cut (the reference fails at reflect.py:790), multiple_reflect (find_dz adds the heights to the NORMAL in
    the derivative stage, base.py:823-845)."""
    from xrt_b200 import figure_error as pfe, oes, materials
    from xrt_b200.beamline import BeamLine
    fe = pfe.Waviness(amplitude=5., xWaveLength=1., yWaveLength=2., limPhysX=[-1, 1], limPhysY=[-2, 2], gridStep=0.25)
    si = materials.CrystalSi(hkl=(1, 1, 1))
    bl = BeamLine()
    lim = dict(limPhysX=[-1, 1], limPhysY=[-2, 2])
    dcm = oes.DCM(bl, 'dcm', center=(0, 1e4, 0), bragg=0.2, material=si, figureError=fe, **lim)
    assert dcm._plugin().kind == 0 and dcm._plugin().figure_error
    pl = oes.Plate(bl, 'w', center=(0, 1e4, 0), pitch=1., t=0.1, figureError=fe,
                   material=materials.Material('C', rho=3.52, kind='plate'), **lim)
    assert pl._plugin() is dcm._plugin()  # one object per (surface kind, figure error)
    ln = oes.ParaboloidFlatLens(bl, 'l', center=(0, 1e4, 0), t=0.05, focus=0.2, zmax=0.5, figureError=fe,
                                material=materials.Material('Be', rho=1.848, kind='lens'), **lim)
    assert ln._plugin().kind == 3 and ln._plugin().figure_error
    asym = oes.DCM(bl, 'dcm', center=(0, 1e4, 0), bragg=0.2, material=si, alpha=0.05, figureError=fe, **lim)
    with pytest.raises(NotImplementedError):
        asym._plugin()
    m = oes.OE(bl, 'm', center=(0, 1e4, 0), pitch=3e-3, figureError=fe, **lim)
    with pytest.raises(NotImplementedError):
        m.multiple_reflect(None)
    assert oes.OE(bl, 'm0', center=(0, 1e4, 0), pitch=3e-3, **lim)._plugin() is None
