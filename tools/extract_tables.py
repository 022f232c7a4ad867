#!/usr/bin/env python
"""Extract the atomic scattering-factor tables the hot path needs into xrt_b200/data/elements.npz.

The f1/f2 (Chantler/NIST FFAST), f0 (Waasmaier-Kirfel / XOP) and atomic-mass tables are physical data the
reference ships under xrt/backends/raycing/data/.  The hot-path configurations need only a handful of mirror
coatings and crystal elements, so this script (run once in the build container, where /root/reference is
mounted) dumps those elements verbatim (float32, as stored) for the product to load at run time.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, '/root/reference')
sys.path.insert(0, os.path.join(HERE, 'ref_stubs'))
import numpy as np  # noqa: E402
import xrt.backends.raycing.materials as rm  # noqa: E402

ELEMENTS = ['Be', 'C', 'O', 'Al', 'Si', 'Cr', 'Ni', 'Cu', 'Ge', 'Mo', 'Rh', 'Pd', 'Ag', 'W', 'Ir', 'Pt', 'Au']
out = {}
for name in ELEMENTS:
    e = rm.Element(name, table='Chantler')
    out[name + '_Z'] = np.int32(e.Z)
    out[name + '_mass'] = np.float64(e.mass)
    out[name + '_f0'] = np.array(e.f0coeffs, dtype=np.float64)
    out[name + '_E'] = np.array(e.E)
    out[name + '_f1'] = np.array(e.f1)
    out[name + '_f2'] = np.array(e.f2)
    et = rm.Element(name, table='Chantler total')  # the default table of Material (material.py:26)
    assert np.array_equal(et.E, e.E) and np.array_equal(et.f1, e.f1)
    out[name + '_f2tot'] = np.array(et.f2)
    assert e.E.dtype == np.float32
path = os.path.join(ROOT, 'xrt_b200', 'data', 'elements.npz')
np.savez_compressed(path, table=np.array('Chantler'), **out)
print('wrote', path, os.path.getsize(path))
