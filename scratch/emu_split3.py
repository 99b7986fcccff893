"""scratch: traversal work (CPU emulation) with and without leaf splitting."""
import os, sys, time; sys.path.insert(0,'/root/repo/scratch')
from emu3 import *
import emu3 as emu2
R = float(sys.argv[1]) if len(sys.argv) > 1 else 6000.0
specs = [tuple(float(x) for x in a.split(',')) for a in sys.argv[2:]] or [(0,8,2.0),(4,8,2.0),(8,8,2.0),(16,8,2.0)]
geo = demo.detector(pmt_radius=R, sphere_radius=R+500.0, spiral_step=350.0); geo.flatten(dedupe_vertices=False)
print('triangles', len(geo.mesh.triangles), 'pmts', geo.num_channels())
scenes.with_bvh(geo)
rng = np.random.default_rng(1); n=20000
d = uniform_sphere(n, rng=rng).astype(np.float32); o = np.zeros((n,3),np.float32)
o2 = (rng.uniform(-1,1,(n,3))*R/2).astype(np.float32)
# rays starting on PMT surfaces (like later propagation steps): from random hit points, random directions
desc, keep = scenes.desc_of(geo)
solid = np.ascontiguousarray(geo.solid_id, dtype=np.uint32)
ref_nodes = keep['nodes']
base = None
for mp, me, mr in specs:
    t=time.time()
    nat = np.ascontiguousarray(native_tree(ref_nodes, len(geo.mesh.triangles), None if os.environ.get('SINGLE') else geo.solid_id, mesh=geo.mesh, world_coords=geo.bvh.world_coords, max_pieces=int(mp), min_extent=int(me), min_ratio=mr))
    bt=time.time()-t
    nleaf = int(((nat['w']>>28)==0).sum())
    desc.nodes = nat.ctypes.data; desc.nnodes = len(nat)
    print('--- max_pieces %d min_extent %d ratio %.1f: build %.1fs entries %d (%.2fx), leaf entries/triangle %.2f' % (mp, me, mr, bt, len(nat), len(nat)/ (base or len(nat)), nleaf/len(geo.mesh.triangles)))
    if base is None: base = len(nat)
    t0, c = run(desc, solid, o, d); report('  from centre', c)
    w = solid[np.maximum(t0,0)] != 0; per = emu2.per
    print('     pmt-winner rays: tris %.1f rounds %.1f | grazing liner rays (%.3f): tris %.1f' % (per[w,1].mean(), per[w,2].mean(), (per[~w,1]>0).mean(), per[~w,1][per[~w,1]>0].mean()))
    if 'ref0' not in globals(): ref0 = t0
    t1, c = run(desc, solid, o2, d); report('  random origins', c)
    if 'ref1' not in globals(): ref1 = t1
    print('     agree with unsplit: %.5f %.5f' % ((t0==ref0).mean(), (t1==ref1).mean()))
