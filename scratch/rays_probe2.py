"""Scratch: why cb_intersect on the 1.2 M-triangle scene takes 2.3 ms in one process and 6-10 ms in another:
per-call times without / with the L2 flush, for three GPUGeometry instances of the same scene in one process."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import bench
from chroma_lite_b200 import gpu, _lib
from chroma_lite_b200 import gpuarray as ga
from chroma_lite_b200.gpu.tools import to_float3
from chroma_lite_b200.bvh import make_recursive_grid_bvh
_lib.init(0); lib = _lib.lib()
geo = bench.rays_scene(); geo.bvh = make_recursive_grid_bvh(geo.mesh)
o, d = bench.make_rays(geo, 10000000)
do, dd = ga.to_gpu(to_float3(o)), ga.to_gpu(to_float3(d))

def timed(g, flush):
    if flush:
        lib.cb_flush_l2()
    lib.cb_synchronize(); lib.cb_timer_start()
    tri, dist = gpu.intersect_mesh(g, do, dd)
    t = _lib.C.c_float(); lib.cb_timer_stop(_lib.C.byref(t))
    return t.value, tri

ref = None
for inst in range(3):
    for window in ('16', '0'):
        os.environ['CHROMA_B200_L2_WINDOW_MB'] = window
        g = gpu.GPUGeometry(geo)
        a = [timed(g, False)[0] for _ in range(40)]
        b = [timed(g, True)[0] for _ in range(12)]
        tri = timed(g, False)[1].get()
        ref = tri if ref is None else ref
        print('instance %d L2 window %s MB: no flush first %s ... last %s | with flush %s | identical %s' % (
            inst, window, ' '.join('%.2f' % x for x in a[:6]), ' '.join('%.2f' % x for x in a[-4:]), ' '.join('%.2f' % x for x in b),
            bool((tri == ref).all())), flush=True)
        os.environ['CHROMA_B200_STATS'] = '1'
        timed(g, False)
        os.environ.pop('CHROMA_B200_STATS')
        del g
