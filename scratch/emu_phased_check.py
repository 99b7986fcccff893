"""scratch: the engine's plane test (hit_box_phased arithmetic, restated in scratch/emu5.c) on split and single-level
trees: nearest-hit distances must equal those of the reference tree walked with the reference's exact box test, for
random rays and for rays aimed at triangle corners / edge midpoints."""
import os, sys, ctypes as C; sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo')
import numpy as np, scenes
from chroma_lite_b200.bvh import native_tree
from chroma_lite_b200.sample import uniform_sphere
from chroma_lite_b200 import demo
emu = C.CDLL('/root/repo/scratch/libemu5.so')
def run(desc, solid, o, d, phased):
    emu.emu2_set_widen(C.c_float(0.0)); emu.emu2_set_phased(C.c_int(phased)); emu.emu2_set_leaf_mode(C.c_int(2 if phased else 0))
    n=len(o); tri=np.full(n,-1,np.int32); cnt=np.zeros(8,np.uint64); per=np.zeros((n,3),np.uint16); dist=np.full(n,-1,np.float32)
    emu.emu2_set_dist_out(dist.ctypes.data_as(C.c_void_p))
    emu.emu2_intersect(C.byref(desc), solid.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p), C.c_uint64(n), tri.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p), per.ctypes.data_as(C.c_void_p))
    emu.emu2_set_dist_out(None)
    return tri, dist, cnt.astype(float)/n
def hit_dist(geo, o, d, tri):
    v = geo.mesh.assemble()[np.maximum(tri,0)].astype(np.float64)
    n = np.cross(v[:,1]-v[:,0], v[:,2]-v[:,0]); dn = d/np.linalg.norm(d,axis=1)[:,None]
    t = np.einsum('ij,ij->i', v[:,0]-o, n)/np.einsum('ij,ij->i', dn, n)
    return np.where(tri>=0, t, -1.0)
cases = [('tiny', scenes.tiny_detector()), ('scint', scenes.scintillator_scene(12)), ('sphere', scenes.sphere_scene(32))]
g = demo.detector(pmt_radius=23775.0, sphere_radius=24275.0, spiral_step=350.0, max_pmts=300); g.flatten(dedupe_vertices=False); scenes.with_bvh(g); cases.append(('cap300 (full-size world grid)', g))
for name, geo in cases:
    rng = np.random.default_rng(11); n = 60000
    lo, hi = geo.mesh.get_bounds()
    o = ((lo+hi)/2 + rng.uniform(-0.55,0.55,(n,3))*(hi-lo)).astype(np.float32); d = uniform_sphere(n, rng=rng).astype(np.float32)
    verts = geo.mesh.assemble(); pick = rng.integers(0, len(verts), 20000)
    aim = np.concatenate([verts[pick, rng.integers(0,3,20000)], (verts[pick,0]+verts[pick,1])/2]).astype(np.float32)
    src = ((lo+hi)/2 + 0.01*(hi-lo)).astype(np.float32)
    o = np.concatenate([o, np.tile(src, (len(aim),1))]); d = np.concatenate([d, aim-src]).astype(np.float32)
    desc, keep = scenes.desc_of(geo); solid = np.ascontiguousarray(geo.solid_id if geo.solid_id is not None else np.zeros(len(geo.mesh.triangles)), dtype=np.uint32)
    ref_tri, ref_t, _ = run(desc, solid, o, d, 0)                      # reference tree, reference box test
    for label, sid, mp in (('solids first', geo.solid_id, 0), ('solids first + split', geo.solid_id, 8), ('single + split', None, 8), ('single + split 32/1.05', None, 32)):
        nat = np.ascontiguousarray(native_tree(keep['nodes'], len(geo.mesh.triangles), sid, mesh=geo.mesh, world_coords=geo.bvh.world_coords, max_pieces=mp, min_extent=4, min_ratio=1.05 if mp == 32 else 2.0))
        desc.nodes = nat.ctypes.data; desc.nnodes = len(nat)
        tri, t, c = run(desc, solid, o, d, 1)                       # engine tree, engine plane test
        same_hit = ((tri >= 0) == (ref_tri >= 0)).all()
        dt = np.abs(t - ref_t)[tri >= 0]; nbit = int((t.view(np.uint32) != ref_t.view(np.uint32)).sum())
        print('%-30s %-24s hit/miss identical %s  same triangle %.5f  distances not bit-identical: %d of %d, max |dt| %.3g  (entries/ray %.1f)' % (name, label, same_hit, (tri == ref_tri).mean(), nbit, len(t), dt.max() if len(dt) else 0, c[1]))
    desc.nodes = keep['nodes'].ctypes.data; desc.nnodes = len(keep['nodes'])
