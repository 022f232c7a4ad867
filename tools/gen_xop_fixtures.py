#!/usr/bin/env python
"""Collect the XOP curves the reference's own test suite compares against (tests/raycing/test_materials.py:239-352
rocking curves, :439-495 mirror reflectivities) into one small fixture, tests/golden/xop_fixtures.npz.

They are outputs of XOP (XCrystal, Xf1f2), not of xrt: an independent physics anchor for the crystal and mirror
amplitudes (the reference only overlays them in plots; tests/test_xop_fixtures.py asserts the distance)."""
import os

import numpy as np

REF = '/root/reference/tests/raycing'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

out = {}
for alpha in (0, 5, -5):
    for pol in 'sp':
        x, R = np.loadtxt(os.path.join(REF, 'XOP-RockingCurves', 'bSi111_thick_{0:-.0f}_{1}.xc.gz'.format(alpha, pol)),
                          unpack=True)
        out['bSi111_thick_{0}_{1}_x'.format(alpha, pol)] = x
        out['bSi111_thick_{0}_{1}_R'.format(alpha, pol)] = R
for name in ('Si05deg', 'Pt4mrad', 'SiO205deg', 'Rh2mrad'):
    for pol in 'sp':
        x, R = np.loadtxt(os.path.join(REF, 'XOP-Reflectivities', '{0}_{1}.xf1f2.gz'.format(name, pol)), unpack=True)
        out['{0}_{1}_E'.format(name, pol)] = x
        out['{0}_{1}_R'.format(name, pol)] = R
path = os.path.join(ROOT, 'tests', 'golden', 'xop_fixtures.npz')
np.savez_compressed(path, **out)
print('wrote', path, os.path.getsize(path), 'bytes')
