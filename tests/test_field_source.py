"""f3: SourceFromField (sources/synchr.py:612-1347) -- tabulated magnetic field -> RK4 trajectory -> radiation
integral along it (kernel `custom_field`, cl/undulator.cl:1001-...): oracle pinned bit-exact on the reference (CPU),
CUDA kernel vs oracle / golden (GPU)."""
import numpy as np
import pytest

from conftest import golden
from oracle import field_source as ofs

CASES = ['field_source_far', 'field_source_near']


def _spec(g):
    R0 = None if np.isnan(g['R0']) else float(g['R0'])
    return ofs.FieldSpec(g['field'], float(g['gamma0']), float(g['eI']), int(g['quadm']), int(g['gIntervals']),
                         R0=R0)


@pytest.mark.parametrize('name', CASES)
def test_oracle_matches_reference_bit_exact(name):
    g = golden(name)
    spec = _spec(g)
    assert np.array_equal(spec.tg, g['tg']) and np.array_equal(spec.ag, g['ag']) and spec.dstep == float(g['dstep'])
    assert spec.deviceLength == float(g['deviceLength'])
    I, Es, Ep = spec.build_I_map(g['w'], g['theta'], g['psi'], g['gamma'])
    assert np.array_equal(I, g['I'])
    assert np.array_equal(Es, g['Es'])
    assert np.array_equal(Ep, g['Ep'])
    assert I.max() > 0


def _host_source(g):
    """The product class without its GPU-touching constructor: enough for the host-side trajectory code."""
    from xrt_b200.sources import SourceFromField
    src = SourceFromField.__new__(SourceFromField)
    src.customFieldData = g['field']
    src.spl_kw = {'kind': 'cubic', 'bounds_error': False, 'fill_value': 'extrapolate'}
    src.quadm, src.gIntervals = int(g['quadm']), int(g['gIntervals'])
    src._traj = None
    src._build_integration_grid()
    return src


@pytest.mark.parametrize('name', CASES)
def test_host_trajectory_matches_oracle_bit_exact(name):
    g = golden(name)
    src = _host_source(g)
    spec = _spec(g)
    assert np.array_equal(src.tg, spec.tg) and np.array_equal(src.ag, spec.ag) and src.dstep == spec.dstep
    betam, tabs = src._field_tables()
    ref = [spec.tg, spec.ag, spec.Bt[0], spec.Bt[1], spec.Bt[2], spec.traj[0], spec.traj[1], spec.traj[3],
           spec.traj[4], spec.traj[5]]
    assert betam == spec.traj[2][-1]
    for a, b in zip(tabs, ref):
        assert np.array_equal(a, b)
    assert src._fwhm(g['field'][:, 0], np.abs(g['field'][:, 1:]).max(axis=1)) == float(g['deviceLength'])


# Every expression of the kernel is in the reference's rounding order (the phase is a 1e7 cancellation); what is left
# are <= 2 ulp differences of sin / cos at arguments up to 1e12 rad (near field: wc * R0): 1e-11 of sum |terms|.
TOL = 1e-11


@pytest.mark.gpu
@pytest.mark.parametrize('name', CASES)
def test_cuda_run_parallel_custom_field_vs_oracle(name, cuda_device):
    """XRT_CUDA.run_parallel('custom_field', ...) with the argument list of _build_I_map_custom_field_CL
    (synchr.py:1227-1259)."""
    from xrt_b200.cuda_backend import XRT_CUDA
    from oracle.physconsts import SIE0, SIM0, C
    g = golden(name)
    spec = _spec(g)
    betax, betay, betazav, trajx, trajy, trajz = spec.traj
    betam = betazav[-1]
    n = len(g['w'])
    gamma = g['gamma'] * np.ones(n)
    R0 = 0. if spec.R0 is None else spec.R0
    ucl = XRT_CUDA('undulator.cl', targetOpenCL='CUDA:0')
    Is = np.zeros(n, dtype=np.complex128)
    Ip = np.zeros(n, dtype=np.complex128)
    ret = ucl.run_parallel('custom_field', [np.int32(len(spec.tg)), np.float64(betam), np.float64(R0)],
                           [gamma, g['w'], g['theta'], g['psi']],
                           [spec.tg, spec.ag, spec.Bt[0], spec.Bt[1], spec.Bt[2], betax, betay, trajx, trajy, trajz],
                           [Is, Ip], None, n)
    assert ret[0] is Is and ret[1] is Ip
    emcg = SIE0 / SIM0 / C / 10. / gamma
    R0v = None
    if spec.R0:
        R0v = np.array((np.tan(g['theta']), np.tan(g['psi']), np.ones(n))) * spec.R0
    ref_s, ref_p, ss, sp = ofs.sp_sum(spec.tg, spec.ag, emcg, g['w'], gamma, g['theta'], g['psi'], spec.Bt[0],
                                      spec.Bt[1], spec.Bt[2], betax, betay, betam, trajx, trajy, trajz, R0v, True)
    assert (np.abs(Is - ref_s) / ss).max() < TOL
    assert (np.abs(Ip - ref_p) / (ss + sp)).max() < TOL


@pytest.mark.gpu
@pytest.mark.parametrize('name', CASES)
def test_cuda_source_from_field_vs_reference_golden(name, cuda_device):
    import torch
    from xrt_b200.sources import SourceFromField
    g = golden(name)
    spec = _spec(g)
    I0, Es0, Ep0, ss, sp = spec.build_I_map(g['w'], g['theta'], g['psi'], g['gamma'], with_scale=True)
    src = SourceFromField(eE=3., eI=float(g['eI']), eEpsilonX=0.263, eEpsilonZ=0.008, eSigmaX=48.65, eSigmaZ=6.197,
                          xPrimeMax=0.1, zPrimeMax=0.1, eMin=float(g['E_min']), eMax=float(g['E_max']),
                          customField=g['field'], gNodes=int(g['quadm']), gIntervals=int(g['gIntervals']),
                          R0=spec.R0, eEspread=float(g['eEspread']))
    assert src.gamma == float(g['gamma0']) and src.deviceLength == float(g['deviceLength'])
    dg = g['gamma'] - float(g['gamma0'])
    I, Es, Ep = src.build_I_map(g['w'], g['theta'], g['psi'], dg=dg if src.eEspread > 0 else None)
    assert (np.abs(Es - g['Es']) / ss).max() < TOL
    assert (np.abs(Ep - g['Ep']) / (ss + sp)).max() < TOL
    assert (np.abs(I - g['I']) / (ss + sp)**2).max() < TOL
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    Id, Esd, Epd = src.build_I_map_device(dev(g['w']), dev(g['theta']), dev(g['psi']), gamma=dev(g['gamma']))
    assert (np.abs(Esd.cpu().numpy() - g['Es']) / ss).max() < TOL
    assert (np.abs(Id.cpu().numpy() - g['I']) / (ss + sp)**2).max() < TOL
    # and a beam: shine() runs sample -> custom_field -> fill on the device
    src.nrays = 4000
    src.uniformRayDensity = True
    beam = src.shine().to_numpy()
    assert len(beam['x']) == 4000 and np.all(beam['state'] == 1)
    assert np.all(np.isfinite(beam['Jss'])) and beam['Jss'].max() > 0
    assert np.abs(beam['Jss'] - np.abs(beam['Es'])**2).max() <= 1e-14 * beam['Jss'].max()
