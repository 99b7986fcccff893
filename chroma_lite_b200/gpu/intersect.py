"""Ray/mesh nearest-hit query (role of the distance_to_mesh kernel,
chroma/cuda/mesh.h:131-155), also returning the triangle index."""
import numpy as np

from .. import _lib
from .. import gpuarray as ga
from .tools import to_float3


def intersect_mesh(gpu_geometry, origins, directions, last_hit_triangles=None):
    """origins, directions: (N,3) host arrays or float3 DeviceArrays.  Returns
    (triangle int32[N] with -1 = miss, distance float32[N] untouched (0) on miss)."""
    lib = _lib.lib()
    o = origins if isinstance(origins, ga.DeviceArray) else ga.to_gpu(to_float3(np.asarray(origins)))
    d = directions if isinstance(directions, ga.DeviceArray) else ga.to_gpu(to_float3(np.asarray(directions)))
    n = len(o)
    lh = None
    if last_hit_triangles is not None:
        lh = last_hit_triangles if isinstance(last_hit_triangles, ga.DeviceArray) else ga.to_gpu(np.asarray(last_hit_triangles, dtype=np.int32))
    tri = ga.empty(n, np.int32)
    dist = ga.zeros(n, np.float32)
    _lib.check(lib.cb_intersect(gpu_geometry.handle, o.ptr, d.ptr, lh.ptr if lh is not None else None, n, tri.ptr, dist.ptr))
    return tri, dist
