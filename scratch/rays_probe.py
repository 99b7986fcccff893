"""Scratch: per-call times of cb_intersect on the rays workload under env settings given as args (K=V,K=V ...)."""
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
g = gpu.GPUGeometry(geo)
do, dd = ga.to_gpu(to_float3(o)), ga.to_gpu(to_float3(d))
for spec in sys.argv[1:]:
    keys = []
    for kv in filter(None, spec.split(',')):
        k, v = kv.split('='); os.environ['CHROMA_B200_' + k] = v; keys.append('CHROMA_B200_' + k)
    ms = []
    for _ in range(12):
        lib.cb_flush_l2(); lib.cb_synchronize(); lib.cb_timer_start()
        tri, dist = gpu.intersect_mesh(g, do, dd)
        t = _lib.C.c_float(); lib.cb_timer_stop(_lib.C.byref(t)); ms.append(t.value)
    print('%-28s' % spec, ' '.join('%.2f' % m for m in ms), flush=True)
    for k in keys: os.environ.pop(k)
