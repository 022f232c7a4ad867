"""a2-a4: undulator grid + far-field integral -- oracle pinned on reference outputs (CPU), CUDA vs oracle (GPU)."""
import numpy as np
import pytest

from conftest import golden
from oracle import undulator as ou

CASES = ['undulator_planar_h1', 'undulator_helical_spread']


def _grid(g):
    return ou.build_integration_grid(int(g['quadm']), int(g['gIntervals']), float(g['phase']))


@pytest.mark.parametrize('name', CASES)
def test_oracle_matches_reference_bit_exact(name):
    g = golden(name)
    grid = _grid(g)
    for k in ('tg', 'ag', 'sintg', 'costg', 'sintgph', 'costgph'):
        assert np.array_equal(grid[k], g[k]), k
    assert grid['dstep'] == float(g['dstep'])
    assert ou.gamma_from_eE(3.) == float(g['gamma0'])
    I, Es, Ep = ou.build_I_map(grid, float(g['L0']), float(g['Np']), float(g['Kx']), float(g['Ky']),
                               float(g['eI']), g['gamma'], g['w'], g['theta'], g['psi'])
    assert np.array_equal(I, g['I'])
    assert np.array_equal(Es, g['Es'])
    assert np.array_equal(Ep, g['Ep'])


def _term_scale(grid, g):
    """sum_j |term_j| of the node sum for Es/Ep (the honest scale of an oscillatory quadrature)."""
    Kx, Ky = float(g['Kx']), float(g['Ky'])
    gam = g['gamma']
    wu, ww1, ab = ou.sample_prefactors(float(g['L0']), float(g['Np']), Kx, Ky, gam, g['w'], g['theta'], g['psi'])
    revg = 1. / gam
    ss = np.zeros(len(gam))
    sp = np.zeros(len(gam))
    th, ps = g['theta'], g['psi']
    dirz = 1. - 0.5 * (th**2 + ps**2)
    for i in range(len(grid['tg'])):
        bx = Ky * revg * grid['costg'][i]
        by = -Kx * revg * grid['costgph'][i]
        bz = 1. - 0.5 * (revg**2 + bx * bx + by * by)
        bpx = -Ky * grid['sintg'][i]
        bpy = Kx * grid['sintgph'][i]
        bpz = 0.5 * revg * (Ky**2 * 2 * grid['sintg'][i] * grid['costg'][i] +
                            Kx**2 * 2 * grid['sintgph'][i] * grid['costgph'][i])
        rk = 1. / (1. - th * bx - ps * by - dirz * bz)
        wgt = np.abs(grid['ag'][i]) * rk**2
        d1 = np.abs(th * bpx) + np.abs(ps * bpy) + np.abs(dirz * bpz)
        d2 = np.abs(th * (th - bx)) + np.abs(ps * (ps - by)) + np.abs(dirz * (dirz - bz))
        ss += wgt * (np.abs(th - bx) * d1 + np.abs(bpx) * d2)
        sp += wgt * (np.abs(ps - by) * d1 + np.abs(bpy) * d2)
    return wu * revg * ss, wu * revg * sp, wu, ww1, ab


# Tolerance: with the cancellation-prone chain evaluated in the reference's rounding order the remaining
# differences are ~1e-14 rad in the phase and a few ulps in the weights: assert 1e-12 * sum|terms| (SURVEY 8c
# asks for rel 1e-10).
TOL = 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize('name', CASES)
def test_cuda_run_parallel_undulator_vs_golden(name, cuda_device):
    """XRT_CUDA.run_parallel('undulator', ...) called the way Undulator._build_I_map_CL calls XRT_CL."""
    from xrt_b200.cuda_backend import XRT_CUDA
    g = golden(name)
    grid = _grid(g)
    ss, sp, wu, ww1, ab = _term_scale(grid, g)
    ucl = XRT_CUDA('undulator.cl', targetOpenCL='CUDA:0')
    n = len(g['w'])
    Is = np.zeros(n, dtype=np.complex128)
    Ip = np.zeros(n, dtype=np.complex128)
    ret = ucl.run_parallel(
        'undulator', [np.float64(0.), np.float64(g['Kx']), np.float64(g['Ky']), np.int32(len(grid['tg']))],
        [g['gamma'], wu, g['w'], ww1, g['theta'], g['psi']],
        [grid[k] for k in ('tg', 'ag', 'sintg', 'costg', 'sintgph', 'costgph')], [Is, Ip], dimension=n)
    assert ret[0] is Is and ret[1] is Ip
    ref_s, ref_p = ou.sp_sum(grid, float(g['Kx']), float(g['Ky']), ww1, g['w'], wu, g['gamma'], g['theta'], g['psi'])
    assert (np.abs(Is - ref_s) / ss).max() < TOL
    if float(g['Kx']) != 0:
        assert (np.abs(Ip - ref_p) / sp).max() < TOL
    else:
        assert (np.abs(Ip - ref_p) / (sp + ss)).max() < TOL
    # and against the reference's own final amplitudes
    from oracle.physconsts import FINE_STR, SIE0
    amp = np.sqrt(FINE_STR / g['w'] * float(g['eI']) / SIE0) * ab * 0.5 * float(g['dstep'])
    assert (np.abs(amp * Is - g['Es']) / np.abs(amp * ss)).max() < TOL


@pytest.mark.gpu
@pytest.mark.parametrize('name', CASES)
@pytest.mark.parametrize('R', [1, 2, 4])
@pytest.mark.parametrize('symmetric', [False, True])
def test_cuda_fused_imap_vs_golden(name, R, symmetric, cuda_device):
    """symmetric=True: the node-pair path for planar undulators on mirror-symmetric grids (ignored by the kernel
    for the helical case, Kx != 0)."""
    import torch
    from xrt_b200 import kernels
    from xrt_b200.physconsts import FINE_STR, SIE0
    g = golden(name)
    grid = _grid(g)
    ss, sp, wu, ww1, ab = _term_scale(grid, g)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    tables = [dev(grid[k]) for k in ('tg', 'ag', 'sintg', 'costg', 'sintgph', 'costgph')]
    I, Es, Ep = kernels.undulator_imap(
        float(g['Kx']), float(g['Ky']), float(g['L0']), float(g['Np']), float(g['gamma0']),
        FINE_STR * float(g['eI']) / SIE0, float(g['dstep']), 0, dev(g['w']), dev(g['theta']), dev(g['psi']),
        tables, gamma=dev(g['gamma']), R=R, symmetric=symmetric)
    amp = np.abs(np.sqrt(FINE_STR / g['w'] * float(g['eI']) / SIE0) * ab * 0.5 * float(g['dstep']))
    # `ab` = sin(pi N ww1)/sin(pi ww1) amplifies 1-ulp differences of ww1 near its zeros: allow 1e-10 of the scale
    assert (np.abs(Es.cpu().numpy() - g['Es']) / (amp * ss + 1e-300)).max() < 1e-10
    assert (np.abs(Ep.cpu().numpy() - g['Ep']) / (amp * (ss + sp) + 1e-300)).max() < 1e-10
    Iref = g['I']
    assert (np.abs(I.cpu().numpy() - Iref) / (amp**2 * (ss + sp)**2)).max() < 1e-10


@pytest.mark.gpu
def test_cuda_many_nodes_tiles(cuda_device):
    """More nodes than one shared-memory tile (256) and a ragged sample count."""
    import torch
    from xrt_b200 import kernels
    g = golden('undulator_planar_h1')
    grid = ou.build_integration_grid(150, 5, 0.3)  # 750 nodes
    n = 777
    sl = slice(0, n)
    Kx, Ky = 0.4, float(g['Ky'])
    wu, ww1, ab = ou.sample_prefactors(float(g['L0']), float(g['Np']), Kx, Ky, g['gamma'][sl], g['w'][sl],
                                       g['theta'][sl], g['psi'][sl])
    ref_s, ref_p = ou.sp_sum(grid, Kx, Ky, ww1, g['w'][sl], wu, g['gamma'][sl], g['theta'][sl], g['psi'][sl])
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    tables = [dev(grid[k]) for k in ('tg', 'ag', 'sintg', 'costg', 'sintgph', 'costgph')]
    Is, Ip = kernels.undulator_raw(Kx, Ky, dev(g['gamma'][sl]), dev(wu), dev(g['w'][sl]), dev(ww1),
                                   dev(g['theta'][sl]), dev(g['psi'][sl]), tables)
    scale = max(np.abs(ref_s).max(), np.abs(ref_p).max())
    assert np.abs(Is.cpu().numpy() - ref_s).max() < 1e-9 * scale
    assert np.abs(Ip.cpu().numpy() - ref_p).max() < 1e-9 * scale


@pytest.mark.gpu
@pytest.mark.parametrize('quadm,gI', [(33, 2), (31, 3), (200, 4)])
def test_cuda_symmetric_pairs_vs_oracle_odd_and_tiled_grids(quadm, gI, cuda_device):
    """Node-pair path on odd node counts (middle node) and on more than one shared-memory tile."""
    import torch
    from xrt_b200 import kernels
    g = golden('undulator_planar_h1')
    grid = ou.build_integration_grid(quadm, gI, 0.)
    assert kernels.grid_is_symmetric(grid['tg'], grid['ag'], grid['sintg'], grid['costg'])
    n = 500
    sl = slice(0, n)
    Kx, Ky = 0., float(g['Ky'])
    wu, ww1, ab = ou.sample_prefactors(float(g['L0']), float(g['Np']), Kx, Ky, g['gamma'][sl], g['w'][sl],
                                       g['theta'][sl], g['psi'][sl])
    ref_s, ref_p = ou.sp_sum(grid, Kx, Ky, ww1, g['w'][sl], wu, g['gamma'][sl], g['theta'][sl], g['psi'][sl])
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    tables = [dev(grid[k]) for k in ('tg', 'ag', 'sintg', 'costg', 'sintgph', 'costgph')]
    args = (Kx, Ky, dev(g['gamma'][sl]), dev(wu), dev(g['w'][sl]), dev(ww1), dev(g['theta'][sl]), dev(g['psi'][sl]),
            tables)
    Is, Ip = kernels.undulator_raw(*args, symmetric=True)
    Is0, Ip0 = kernels.undulator_raw(*args, symmetric=False)
    scale = max(np.abs(ref_s).max(), np.abs(ref_p).max())
    for a, b in ((Is, ref_s), (Ip, ref_p), (Is0, ref_s), (Ip0, ref_p)):
        assert np.abs(a.cpu().numpy() - b).max() < 1e-9 * scale
    assert float((Is - Is0).abs().max()) < 1e-11 * scale


# ------------------------------------------------------------------ f3: tapered and near-field undulators
EXT_CASES = ['undulator_taper', 'undulator_taper_elliptical', 'undulator_nf', 'undulator_nf_elliptical']


def _ext_args(g):
    taper = None if np.isnan(g['taperVal']) else float(g['taperVal'])
    R0 = None if np.isnan(g['R0']) else float(g['R0'])
    return taper, R0


@pytest.mark.parametrize('name', EXT_CASES)
def test_oracle_ext_matches_reference_bit_exact(name):
    g = golden(name)
    grid = _grid(g)
    for k in ('tg', 'ag', 'sintg', 'costg', 'sintgph', 'costgph'):
        assert np.array_equal(grid[k], g[k]), k
    taper, R0 = _ext_args(g)
    assert (taper is None) != (R0 is None)
    I, Es, Ep = ou.build_I_map(grid, float(g['L0']), int(g['Np']), float(g['Kx']), float(g['Ky']),
                               float(g['eI']), g['gamma'], g['w'], g['theta'], g['psi'], R0=R0, taperVal=taper)
    assert np.array_equal(I, g['I'])
    assert np.array_equal(Es, g['Es'])
    assert np.array_equal(Ep, g['Ep'])


def _ext_reference_sums(g, grid):
    taper, R0 = _ext_args(g)
    L0, Np, Kx, Ky = float(g['L0']), int(g['Np']), float(g['Kx']), float(g['Ky'])
    wu, ww1, _ = ou.sample_prefactors(L0, Np, Kx, Ky, g['gamma'], g['w'], g['theta'], g['psi'])
    R0v = None
    scalar0 = taper
    if R0 is not None:
        R0v = np.array((np.tan(g['theta']), np.tan(g['psi']), np.ones_like(g['psi'])))
        R0v *= R0*np.pi*2/L0
        scalar0 = R0*np.pi*2/L0
    ref_s, ref_p, ss, sp = ou.sp_sum(grid, Kx, Ky, ww1, g['w'], wu, g['gamma'], g['theta'], g['psi'], R0v, taper, Np,
                                     with_scale=True)
    return wu, ww1, scalar0, ref_s, ref_p, ss, sp


# Tolerance: the kernel evaluates the reference's expressions in the reference's rounding order (-fmad=false,
# IEEE sqrt and division); what remains are <= 2 ulp differences of sin/cos.  In the near field the phase
# k*(drs + ...) is ~1e3..1e4 rad, one ulp of which is ~1e-12 rad: assert 1e-11 * sum|terms|.
EXT_TOL = 1e-11


@pytest.mark.gpu
@pytest.mark.parametrize('name', EXT_CASES)
def test_cuda_run_parallel_undulator_ext_vs_oracle(name, cuda_device):
    """XRT_CUDA.run_parallel('undulator_taper' | 'undulator_nf', ...) with the argument list of
    Undulator._build_I_map_CL (synchr.py:2134-2176)."""
    from xrt_b200.cuda_backend import XRT_CUDA
    g = golden(name)
    grid = _grid(g)
    wu, ww1, scalar0, ref_s, ref_p, ss, sp = _ext_reference_sums(g, grid)
    taper, R0 = _ext_args(g)
    ucl = XRT_CUDA('undulator.cl', targetOpenCL='CUDA:0')
    n = len(g['w'])
    Is = np.zeros(n, dtype=np.complex128)
    Ip = np.zeros(n, dtype=np.complex128)
    ret = ucl.run_parallel(
        'undulator_taper' if taper is not None else 'undulator_nf',
        [np.float64(scalar0), np.float64(g['Kx']), np.float64(g['Ky']), np.int32(len(grid['tg'])),
         np.int32(g['Np'])],
        [g['gamma'], wu, g['w'], ww1, g['theta'], g['psi']],
        [grid[k] for k in ('tg', 'ag', 'sintg', 'costg', 'sintgph', 'costgph')], [Is, Ip], dimension=n)
    assert ret[0] is Is and ret[1] is Ip
    assert (np.abs(Is - ref_s) / ss).max() < EXT_TOL
    assert (np.abs(Ip - ref_p) / (sp + ss)).max() < EXT_TOL


@pytest.mark.gpu
@pytest.mark.parametrize('name', EXT_CASES)
def test_cuda_undulator_ext_build_I_map_vs_golden(name, cuda_device):
    """Undulator(taper=... | R0=...).build_I_map (host arrays) and build_I_map_device (fused) against the
    reference's own (I, Es, Ep)."""
    import torch
    from xrt_b200.sources import Undulator
    from oracle.physconsts import FINE_STR, SIE0
    g = golden(name)
    grid = _grid(g)
    wu, ww1, scalar0, ref_s, ref_p, ss, sp = _ext_reference_sums(g, grid)
    taper, R0 = _ext_args(g)
    src = Undulator(eE=3., eI=float(g['eI']), period=float(g['L0']), n=int(g['Np']), K=0, Kx=float(g['Kx']),
                    Ky=float(g['Ky']), phaseDeg=float(np.degrees(g['phase'])), eMin=float(g['E_min']),
                    eMax=float(g['E_max']), gNodes=int(g['quadm']), gIntervals=int(g['gIntervals']),
                    taper=taper, R0=R0, eEspread=float(g['eEspread']), xPrimeMaxAutoReduce=False,
                    zPrimeMaxAutoReduce=False)
    assert src.gamma == float(g['gamma0'])
    dg = g['gamma'] - float(g['gamma0'])
    I, Es, Ep = src.build_I_map(g['w'], g['theta'], g['psi'], dg=dg if src.eEspread > 0 else None)
    amp = np.abs(np.sqrt(FINE_STR / g['w'] * float(g['eI']) / SIE0) / (2 * np.pi) / wu * 0.5 * float(g['dstep']))
    assert (np.abs(Es - g['Es']) / (amp * ss)).max() < EXT_TOL
    assert (np.abs(Ep - g['Ep']) / (amp * (ss + sp))).max() < EXT_TOL
    assert (np.abs(I - g['I']) / (amp**2 * (ss + sp)**2)).max() < EXT_TOL
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    Id, Esd, Epd = src.build_I_map_device(dev(g['w']), dev(g['theta']), dev(g['psi']), gamma=dev(g['gamma']))
    assert (np.abs(Esd.cpu().numpy() - g['Es']) / (amp * ss)).max() < EXT_TOL
    assert (np.abs(Epd.cpu().numpy() - g['Ep']) / (amp * (ss + sp))).max() < EXT_TOL
    assert (np.abs(Id.cpu().numpy() - g['I']) / (amp**2 * (ss + sp)**2)).max() < EXT_TOL


@pytest.mark.gpu
def test_cuda_undulator_taper_shine(cuda_device):
    """shine() of a tapered undulator runs the unfused device path (sample -> undulator_ext -> fill) and its
    flux map equals build_I_map_device on the same samples."""
    import torch
    from xrt_b200.sources import Undulator
    src = Undulator(eE=3., eI=0.5, period=30., n=20, K=1.45, eMin=1400., eMax=1500., gNodes=16, gIntervals=2,
                    taper=(0.4, 10.), uniformRayDensity=True, nrays=5000, xPrimeMax=0.1, zPrimeMax=0.1)
    beam = src.shine()
    assert len(beam) == 5000
    J = (beam.Jss + beam.Jpp).cpu().numpy()
    assert np.all(np.isfinite(J)) and J.max() > 0
    assert int((beam.state == 1).sum()) == 5000


# ------------------------------------------------------------------ flux maps: intensities_on_mesh
MESH_CASES = ['mesh_planar_harmonics', 'mesh_elliptical']


def _mesh_kwargs(g):
    kw = dict(eSpreadSigmas=float(g['eSpreadSigmas']), eSpreadNSamples=int(g['eSpreadNSamples']))
    harm = list(g['harmonic']) if 'harmonic' in g.files else None
    return harm, kw


@pytest.mark.parametrize('name', MESH_CASES)
def test_oracle_intensities_on_mesh_matches_reference_bit_exact(name):
    g = golden(name)
    grid = _grid(g)
    harm, kw = _mesh_kwargs(g)
    ss = ou.intensities_on_mesh(grid, float(g['L0']), float(g['Np']), float(g['Kx']), float(g['Ky']), float(g['eI']),
                                float(g['gamma0']), float(g['eEspread']), float(g['dxprime']), float(g['dzprime']),
                                g['energy'], g['theta'], g['psi'], harm, distE=str(g['distE']), **kw)
    for i in range(4):
        assert np.array_equal(ss[i], g['s%d' % i], equal_nan=True), i
    assert g['s0'].max() > 0


@pytest.mark.gpu
@pytest.mark.parametrize('name', MESH_CASES)
def test_cuda_intensities_on_mesh_vs_reference_golden(name, cuda_device):
    """Undulator.intensities_on_mesh (sybase.py:740-928): Stokes parameters on an (E, theta, psi[, harmonic]) mesh,
    averaged over the energy spread and convolved with the electron divergence like the reference."""
    from xrt_b200.sources import Undulator
    g = golden(name)
    harm, kw = _mesh_kwargs(g)
    zero_emittance = float(g['dxprime']) == 0
    src = Undulator(eE=3., eI=float(g['eI']), period=float(g['L0']), n=int(g['Np']), K=0, Kx=float(g['Kx']),
                    Ky=float(g['Ky']), phaseDeg=float(np.degrees(g['phase'])), eMin=float(g['energy'][0]),
                    eMax=float(g['energy'][-1]), gNodes=int(g['quadm']), gIntervals=int(g['gIntervals']),
                    eEspread=float(g['eEspread']), distE=str(g['distE']), xPrimeMaxAutoReduce=False,
                    zPrimeMaxAutoReduce=False, eSigmaX=0. if zero_emittance else 48.65,
                    eSigmaZ=0. if zero_emittance else 6.197, eEpsilonX=0. if zero_emittance else 0.263,
                    eEpsilonZ=0. if zero_emittance else 0.008)
    assert src.dxprime == float(g['dxprime']) and src.dzprime == float(g['dzprime'])
    ss = src.intensities_on_mesh(g['energy'], g['theta'], g['psi'], harm, **kw)
    s0 = g['s0']
    # flux density: 1e-10 of the map's maximum (the harmonic window zeroes samples at ww1 = h +- 0.5 exactly as the
    # reference does: ww1 is computed on the host in both); degrees of polarisation where there is light
    assert np.abs(ss[0] - s0).max() <= 1e-10 * s0.max()
    lit = s0 > 1e-6 * s0.max()
    for i in (1, 2, 3):
        assert np.abs(ss[i] - g['s%d' % i])[lit].max() <= 1e-8, i


@pytest.mark.gpu
def test_cuda_tuning_curve_follows_the_undulator_equation(cuda_device):
    """Undulator.tuning_curves (synchr.py:1710-1750): the flux maximum of harmonic h sits just below
    h * E1(K) = h * 2 gamma^2 hc / (L0 (1 + K^2 / 2)), for every K."""
    from xrt_b200.sources import Undulator
    src = Undulator(eE=3., eI=0.5, period=30., n=40, K=1.45, eMin=500., eMax=6000., gNodes=40, gIntervals=2,
                    eSigmaX=0., eSigmaZ=0., eEpsilonX=0., eEpsilonZ=0., distE='BW', xPrimeMaxAutoReduce=False,
                    zPrimeMaxAutoReduce=False)
    theta = np.linspace(-1, 1, 9) * 1.5e-5
    psi = np.linspace(-1, 1, 9) * 1.5e-5
    Ks = [0.8, 1.45]
    for K in Ks:
        src.Ky = K
        src.report_E1()
        e1 = src.E1
        energy = np.linspace(0.9 * e1, 1.03 * e1, 40)
        tE, tF = src.tuning_curves(energy, theta, psi, [1], [K])
        assert tE.shape == (1, 1) and tF[0, 0] > 0
        assert 0.97 * e1 < tE[0, 0] * 1e3 <= 1.005 * e1
    p = src.power_vs_K(np.linspace(500., 6000., 12), theta, psi, None, Ks)
    assert p.shape == (2,) and p[1] > p[0] > 0  # more field, more power
