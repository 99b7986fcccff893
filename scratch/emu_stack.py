"""scratch: how many traversal-stack entries the engine's ordered traversal holds (CPU emulation, scratch/emu5.c with the
engine's plane test and LIFO leaf queue), per tree option, on the full 29k-PMT detector.  The kernel keeps CB_PSTACK = 16
entries per lane in shared memory + CB_PLSTACK = 48 in local memory and reports anything beyond as an error
(engine.cuh:316-322), so the maximum here must stay below 64 for an option to be usable at all, and the share above 16
is the traffic of the local-memory spill path.
usage: python scratch/emu_stack.py [pmt_radius [nrays]]   (libemu5.so built as scratch/README.md says)"""
import sys, time, ctypes as C; sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo')
import numpy as np, scenes
from chroma_lite_b200.bvh import native_tree
from chroma_lite_b200.sample import uniform_sphere
from chroma_lite_b200 import demo
emu = C.CDLL('/root/repo/scratch/libemu5.so')


def run(desc, solid, o, d):
    emu.emu2_set_widen(C.c_float(0.0)); emu.emu2_set_phased(C.c_int(1)); emu.emu2_set_leaf_mode(C.c_int(2))
    emu.emu2_reset_occupancy()
    n = len(o); tri = np.full(n, -1, np.int32); cnt = np.zeros(8, np.uint64); per = np.zeros((n, 3), np.uint16)
    emu.emu2_intersect(C.byref(desc), solid.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p),
                       d.ctypes.data_as(C.c_void_p), C.c_uint64(n), tri.ctypes.data_as(C.c_void_p),
                       cnt.ctypes.data_as(C.c_void_p), per.ctypes.data_as(C.c_void_p))
    hist = np.zeros(513, np.uint64)
    mx = emu.emu2_get_occupancy(hist.ctypes.data_as(C.c_void_p))
    return tri, cnt.astype(float) / n, mx, hist.astype(float)


R = float(sys.argv[1]) if len(sys.argv) > 1 else 23775.0
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
t = time.time()
geo = demo.detector(pmt_radius=R, sphere_radius=R + 500.0, spiral_step=350.0); geo.flatten(dedupe_vertices=False)
scenes.with_bvh(geo)
print('detector: %d triangles, %d PMTs, built in %.0f s' % (len(geo.mesh.triangles), geo.num_channels(), time.time() - t), flush=True)
rng = np.random.default_rng(1)
d = uniform_sphere(2 * n, rng=rng).astype(np.float32)
o = np.concatenate([np.zeros((n, 3), np.float32), (rng.uniform(-1, 1, (n, 3)) * R * 0.55).astype(np.float32)])
desc, keep = scenes.desc_of(geo)
solid = np.ascontiguousarray(geo.solid_id, dtype=np.uint32)
ref = None
for label, sid in (('solids first (default)', geo.solid_id), ('single level', None)):
    t = time.time()
    nat = np.ascontiguousarray(native_tree(keep['nodes'], len(geo.mesh.triangles), sid))
    desc.nodes = nat.ctypes.data; desc.nnodes = len(nat)
    tri, c, mx, hist = run(desc, solid, o, d)
    ref = tri if ref is None else ref
    exp = hist.sum()
    print('%-24s build %.0f s, %d entries | expansions/ray %.1f | stack entries held: max %d, mean %.2f, '
          'expansions with > 8: %.4f, > 16 (local-memory spill): %.5f, > 24: %.6f, > 32: %.6f | same triangles as default: %.6f'
          % (label, time.time() - t, len(nat), c[0], mx, (hist * np.arange(513)).sum() / exp, hist[9:].sum() / exp,
             hist[17:].sum() / exp, hist[25:].sum() / exp, hist[33:].sum() / exp, (tri == ref).mean()), flush=True)
