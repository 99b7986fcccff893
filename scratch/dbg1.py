import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, scenes
from chroma_lite_b200 import gpu
from oracle import orc, ref_driver
from test_gpu_intersect import random_rays, run_engine
os.environ['CHROMA_B200_STATS']='1'
geo = scenes.tiny_detector()
n = 400000
o, d = random_rays(geo, n, 1234)
v = geo.mesh.vertices[np.random.default_rng(1).integers(0, len(geo.mesh.vertices), 20000)]
t = geo.mesh.triangles[np.random.default_rng(2).integers(0, len(geo.mesh.triangles), 20000)]
mid = 0.5 * (geo.mesh.vertices[t[:, 0]] + geo.mesh.vertices[t[:, 1]])
src = np.array([3.0, -2.0, 1.0], dtype=np.float32)
o = np.concatenate([o, np.tile(src, (40000, 1))]).astype(np.float32)
d = np.concatenate([d, v - src, mid - src]).astype(np.float32)
tri, dist = run_engine(geo, o, d)
desc, keep = scenes.desc_of(geo)
rg = ref_driver.RefGeometry(desc, keep)
rtri, rdist, _ = ref_driver.intersect(rg, o, d)
bad = np.flatnonzero(tri != rtri)
print("mismatch", len(bad), bad)
otri, odist, cnt = orc.intersect(desc, o[bad], d[bad])
rank = orc.triangle_rank(desc)
for k, i in enumerate(bad):
    print(i, 'mine', tri[i], dist[i], 'ref', rtri[i], rdist[i], 'oracle', otri[k], odist[k], 'ranks', rank[tri[i]] if tri[i]>=0 else None, rank[rtri[i]] if rtri[i]>=0 else None, 'o', o[i], 'd', d[i])
hit = (tri == rtri) & (rtri >= 0)
print('dist mismatch among same tri', (dist[hit].view(np.uint32) != rdist[hit].view(np.uint32)).sum())
