import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from conftest import golden
from helpers import *
import test_reflect as tr
for name in ('reflect_ellipse_vfm','reflect_ellipse_hfm','reflect_ellipsoid'):
    g = golden(name)
    oe = tr._product_oe(g, 'ellipse_param')
    gb, lb = oe.reflect(device_beam(oracle_beam(g,'in_')))
    for pre, b in (('gb_', gb), ('lb_', lb)):
        d = b.to_numpy()
        for f in tr.OUT_FIELDS:
            if pre+f in g.files and f in d:
                if f == 'state':
                    print(name, pre+f, (d[f]!=g[pre+f]).sum())
                else:
                    err = np.abs(d[f]-g[pre+f]).max()
                    if err > 0: print(name, pre+f, 'abs', err, 'scale', np.abs(g[pre+f]).max())
