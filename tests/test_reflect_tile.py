"""The tile structure of the reflect main pass (reflect_tile_kernel: rays ranked by class in shared memory, dense solver /
amplitude warps, bulk-async stores) against reflect_main_kernel, bit for bit, through OE.reflect / DCM.double_reflect ->
xrt_reflect_f64 / xrt_reflect_expose_f64.  The cases live in tools/gpu_tile_check.py (also the script of the GPU sweeps):
toroid + Au with and without the local beam and fused with Screen.expose, beams with dead / over / state-2 rays, ragged
and tiny sizes, a view at an odd offset (plain-copy path), a beam without amplitudes, bare and coated flat mirrors with
optical limits, the elliptic parametric mirror (s, phi, r), the DCM with a small second crystal (global ray taken from
the fallback beam) and an asymmetric cut.  Both kernels are also checked against the reference's goldens by the other
-m gpu tests: the dispatcher sends them through the tile kernel wherever it is built."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_tile_kernel_equals_main_kernel_bitwise(cuda_device, tmp_path):
    env = dict(os.environ)
    env.pop('XRTB_REFLECT_TILE', None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'gpu_tile_check.py'), '--nrays', '300077'],
                       cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = json.loads(r.stdout.strip().split('\n')[-1])
    assert res['all_ok'] and len(res['cases']) >= 29
    assert all(v == 'ok' for k, v in res['cases'].items()), res['cases']


def test_tile_switch_is_read_per_launch():
    """XRTB_REFLECT_TILE=0 must reach the library through the process environment (os.environ -> putenv -> getenv)."""
    src = open(os.path.join(ROOT, 'xrt_b200', 'csrc', 'reflect.cu')).read()
    assert 'getenv("XRTB_REFLECT_TILE")' in src
    assert 'reflect_tile_kernel' in src and 'bulk_s2g' in src
