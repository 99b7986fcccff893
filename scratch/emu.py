import sys, time, ctypes as C; sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo')
import numpy as np, scenes
from chroma_lite_b200.bvh import native_tree, unpack_nodes
from chroma_lite_b200.sample import uniform_sphere
emu = C.CDLL('/root/repo/scratch/libemu.so')
def run(desc, o, d, sort_all):
    n=len(o); tri=np.full(n,-1,np.int32); dist=np.zeros(n,np.float32); cnt=np.zeros(4,np.uint64)
    emu.emu_intersect(C.byref(desc), o.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p), C.c_uint64(n), tri.ctypes.data_as(C.c_void_p), dist.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p), C.c_int(sort_all))
    return tri, dist, cnt.astype(float)/n
def evaluate(name, geo, o, d, nat=None):
    desc, keep = scenes.desc_of(geo)
    res = {}
    for sort_all in (0,1):
        t0,_,c0 = run(desc, o, d, sort_all); res[('ref',sort_all)] = c0
    if nat is None: nat = native_tree(keep['nodes'], len(geo.mesh.triangles), geo.solid_id)
    natc = np.ascontiguousarray(nat); desc.nodes = natc.ctypes.data; desc.nnodes = len(natc)
    for sort_all in (0,1):
        t1,_,c1 = run(desc, o, d, sort_all); res[('nat',sort_all)] = c1
    print(name, 'agree', (t0==t1).mean())
    for k,v in res.items(): print('   %s sort_all=%d: rounds %.1f entries %.1f tris %.2f' % (k[0], k[1], v[0], v[1], v[2]))
if __name__ == '__main__':
    rng = np.random.default_rng(1); n=20000
    for name, geo in (('sphere', scenes.sphere_scene(64)), ('tiny', scenes.tiny_detector())):
        d = uniform_sphere(n, rng=rng).astype(np.float32); o = np.zeros((n,3),np.float32)
        evaluate(name, geo, o, d)
