"""Debug helper (GPU box): where do the stage dumps of the CUDA path and the oracle differ most?"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cases, orc
from grtrans_b200 import _lib, args_from_kwargs
cfg = sys.argv[1] if len(sys.argv) > 1 else "ffjet"
kw = getattr(cases, "cfg_" + cfg)(n=30)
a = args_from_kwargs(**kw)
npix = a.nn[0] * a.nn[1]
g = _lib.ray_stage_dump(a, 1, npix)
d = orc.ray_dump(1, npix, _args=args_from_kwargs(**kw))
for k in ['s2xi', 'c2xi', 'ang', 'rshift', 'cosne']:
    e = np.abs(g[k] - d[k])
    i = np.unravel_index(np.nanargmax(e), e.shape)
    print(k, "max err %.3e at ray %d pt %d of %d" % (e[i], i[0], i[1], d['npts'][i[0]]), "gpu", g[k][i], "orc", d[k][i])
    print("   r %.6f th %.6f  u:" % (d['r'][i], d['th'][i]), g['u'][i] - d['u'][i], " b:", g['b'][i] - d['b'][i], d['b'][i])
    print("   n(err>1e-8) =", int((e > 1e-8).sum()), "of", int((np.arange(e.shape[1])[None, :] < d['npts'][:, None]).sum()))
