"""Scratch: which photons alive after 2 steps turn out to be long-lived? (CPU oracle)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import bench
from chroma_lite_b200.gpu.geometry import make_desc
from chroma_lite_b200 import event
from oracle import orc
det = bench.build_detector('pmt29k', {})
desc, keep = make_desc(det)
n = 400000
ph = bench.make_event(n, seed=999)
st = orc.rng_init(42, 0, n)
b2, _ = orc.propagate(desc, ph, st.copy(), max_steps=2)
f2 = b2.flags.copy(); lh2 = b2.last_hit_triangles.copy()
alive2 = (f2 & event.TERMINAL_MASK) == 0
res = {}
for K in (6, 12, 25):
    bK, _ = orc.propagate(desc, ph, st.copy(), max_steps=K)
    res[K] = (bK.flags & event.TERMINAL_MASK) == 0
print('alive after 2:', alive2.sum())
codes = keep['codes']
for K, aliveK in res.items():
    print('--- alive after %d: %d' % (K, aliveK.sum()))
    for name, mask in (('REFLECT_SPECULAR', (f2 & event.REFLECT_SPECULAR) != 0), ('REFLECT_DIFFUSE', (f2 & event.REFLECT_DIFFUSE) != 0),
                       ('RAYLEIGH', (f2 & event.RAYLEIGH_SCATTER) != 0), ('on_surface(lh>=0)', lh2 >= 0), ('no flags', f2 == 0)):
        m = mask & alive2
        print('  %-20s: %6d of alive2 (%.1f%%) ; contains %5d of the %d long-lived (%.1f%%)' % (
            name, m.sum(), 100.0 * m.sum() / alive2.sum(), (m & aliveK).sum(), aliveK.sum(), 100.0 * (m & aliveK).sum() / max(aliveK.sum(), 1)))
    # surface of the last hit triangle
    lh = lh2[alive2 & aliveK]
    surf = (codes[lh[lh >= 0]] >> 8) & 0xff
    print('  surface codes of last hit among long-lived:', np.unique(surf, return_counts=True))
    lh = lh2[alive2 & (lh2 >= 0)]
    surf = (codes[lh] >> 8) & 0xff
    print('  surface codes of last hit among all alive2:', np.unique(surf, return_counts=True))
