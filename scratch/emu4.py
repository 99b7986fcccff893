import sys, time, ctypes as C; sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo')
import numpy as np, scenes
from chroma_lite_b200.bvh import native_tree
from chroma_lite_b200.sample import uniform_sphere
from chroma_lite_b200 import demo
emu = C.CDLL('/root/repo/scratch/libemu4.so')
import os
emu.emu2_set_widen(C.c_float(float(os.environ.get("WIDEN", "0")))); emu.emu2_set_leaf_mode(C.c_int(int(os.environ.get("LEAFMODE", "0"))))
def run(desc, solid, o, d):
    n=len(o); tri=np.full(n,-1,np.int32); cnt=np.zeros(8,np.uint64); global per; per=np.zeros((n,3),np.uint16)
    emu.emu2_intersect(C.byref(desc), solid.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p), C.c_uint64(n), tri.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p), per.ctypes.data_as(C.c_void_p))
    return tri, cnt.astype(float)/n
def report(name, c):
    print('%-28s rounds %.1f halves %.1f entries %.1f inner_hits %.1f | tris liner %.2f pmt %.2f | winners liner %.3f pmt %.3f | iters~ %.1f' % (name, c[0], c[4], c[1], c[7], c[2], c[3], c[5], c[6], c[4]+c[2]+c[3]))
if __name__ == '__main__':
    R = float(sys.argv[1]) if len(sys.argv) > 1 else 6000.0
    t=time.time(); geo = demo.detector(pmt_radius=R, sphere_radius=R+500.0, spiral_step=350.0); geo.flatten(dedupe_vertices=False); print('flatten', time.time()-t, len(geo.mesh.triangles), geo.num_channels())
    scenes.with_bvh(geo)
    rng = np.random.default_rng(1); n=20000
    d = uniform_sphere(n, rng=rng).astype(np.float32); o = np.zeros((n,3),np.float32)
    desc, keep = scenes.desc_of(geo)
    solid = np.ascontiguousarray(geo.solid_id, dtype=np.uint32)
    nat = np.ascontiguousarray(native_tree(keep['nodes'], len(geo.mesh.triangles), geo.solid_id))
    desc.nodes = nat.ctypes.data; desc.nnodes = len(nat)
    t0, c = run(desc, solid, o, d); report('native, from centre', c)
    w = solid[np.maximum(t0,0)] != 0
    for nm, m in (('winner liner', ~w), ('winner pmt', w)):
        print('  ', nm, 'rays %.3f'%m.mean(), 'tris liner %.2f pmt %.2f rounds %.1f'%(per[m,0].mean(), per[m,1].mean(), per[m,2].mean()), 'pmt-tris quantiles', np.quantile(per[m,1],[.5,.9,.99]), 'rounds q', np.quantile(per[m,2],[.5,.9,.99]))
    print('   liner-winner rays with zero pmt tris: %.3f'%(per[~w,1]==0).mean())
    o2 = (rng.uniform(-1,1,(n,3))*R/2).astype(np.float32)
    t1, c = run(desc, solid, o2, d); report('native, random origins', c)
