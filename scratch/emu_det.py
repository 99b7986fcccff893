import sys, time; sys.path.insert(0,'/root/repo/scratch')
from emu import *
from chroma_lite_b200 import demo
t=time.time(); geo = demo.detector(pmt_radius=6000.0, sphere_radius=6500.0, spiral_step=350.0); geo.flatten(dedupe_vertices=False); print('flatten', time.time()-t, len(geo.mesh.triangles), geo.num_channels())
t=time.time(); scenes.with_bvh(geo); print('bvh', time.time()-t, len(geo.bvh.nodes))
rng = np.random.default_rng(1); n=20000
d = uniform_sphere(n, rng=rng).astype(np.float32); o = np.zeros((n,3),np.float32)
desc, keep = scenes.desc_of(geo)
t=time.time(); nat = native_tree(keep['nodes'], len(geo.mesh.triangles), geo.solid_id); print('native build', time.time()-t, len(nat))
evaluate('det1900 from centre', geo, o, d, nat)
o2 = (rng.uniform(-1,1,(n,3))*3000).astype(np.float32)
evaluate('det1900 random origins', geo, o2, d, nat)
