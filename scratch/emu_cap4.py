"""scratch: traversal work (CPU emulation, slabs widened like the engine's plane test) on a cap of the
29k-PMT detector: same world grid and PMT size in quanta as the full detector, 2,000 PMTs."""
import os, sys, time; sys.path.insert(0,'/root/repo/scratch')
os.environ.setdefault('WIDEN', '1.9')
from emu4 import *
import emu4 as emu3
specs = [tuple(float(x) for x in a.split(',')) for a in sys.argv[1:]] or [(0,8,2.0),(4,8,4.0)]
NP = int(os.environ.get('NPMT', '2000'))
geo = demo.detector(pmt_radius=23775.0, sphere_radius=24275.0, spiral_step=350.0, max_pmts=NP); geo.flatten(dedupe_vertices=False)
print('triangles', len(geo.mesh.triangles), 'pmts', geo.num_channels())
scenes.with_bvh(geo)
pm = np.array(geo.solid_displacements[1:]); axis = pm.mean(axis=0); axis /= np.linalg.norm(axis)
cosmax = (pm @ axis / np.linalg.norm(pm, axis=1)).min()
rng = np.random.default_rng(1); n=20000
def cone(n):
    d = uniform_sphere(8*n, rng=rng); d = d[(d @ axis) > cosmax + 0.02][:n]; return d.astype(np.float32)
d = cone(n); o = np.zeros((len(d),3),np.float32)
# photons re-starting near the PMT cap (later propagation steps): origins 0-3 m in front of the cap, isotropic directions
base = cone(n); o2 = (base * (23775.0 - rng.uniform(100, 3000, (len(base),1)))).astype(np.float32); d2 = uniform_sphere(len(base), rng=rng).astype(np.float32)
desc, keep = scenes.desc_of(geo)
solid = np.ascontiguousarray(geo.solid_id, dtype=np.uint32)
for single in (0, 1):
  for mp, me, mr in specs:
    t=time.time()
    nat = np.ascontiguousarray(native_tree(keep['nodes'], len(geo.mesh.triangles), None if single else geo.solid_id, mesh=geo.mesh, world_coords=geo.bvh.world_coords, max_pieces=int(mp), min_extent=int(me), min_ratio=mr))
    bt=time.time()-t
    nleaf = int(((nat['w']>>28)==0).sum())
    desc.nodes = nat.ctypes.data; desc.nnodes = len(nat)
    print('--- %s, max_pieces %d min_extent %d ratio %.1f: build %.1fs entries %d, leaf entries/triangle %.2f' % ('single-level' if single else 'two-level', mp, me, mr, bt, len(nat), nleaf/len(geo.mesh.triangles)))
    t0, c = run(desc, solid, o, d); report('  centre -> cap', c)
    w = solid[np.maximum(t0,0)] != 0; per = emu3.per
    print('     pmt-winner rays: tris %.1f rounds %.1f | rounds+tris all rays %.1f' % (per[w,1].mean(), per[w,2].mean(), c[0]+c[2]+c[3]))
    t1, c = run(desc, solid, o2, d2); report('  near cap, isotropic', c)
    print('     rounds+tris all rays %.1f' % (c[0]+c[2]+c[3]))
