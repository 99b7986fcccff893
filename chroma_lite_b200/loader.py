"""Geometry loading glue with the interface of chroma/loader.py:131-199: flatten the
object, then take its BVH from the cache or build it (cb_bvh_build on the GPU) and store it.

load_geometry_from_string (chroma/loader.py:13-129: named geometries, STL files, Python
entry points) is the reference's CLI front end and is not part of the transport path.
"""
import logging
import time

from .bvh import make_recursive_grid_bvh
from .cache import Cache, mesh_hash
from .geometry import Geometry, Mesh, Solid, vacuum

logger = logging.getLogger(__name__)


def load_bvh(geometry, bvh_name='default', auto_build_bvh=True, read_bvh_cache=False, update_bvh_cache=True,
             cache_dir=None, cuda_device=None):
    """BVH for the flattened ``geometry``: the cached one when asked for and present, else a new
    recursive-grid tree (saved when ``update_bvh_cache``); None when neither applies
    (chroma/loader.py:131-160)."""
    cache = Cache() if cache_dir is None else Cache(cache_dir)
    key = mesh_hash(geometry.mesh)
    bvh = None
    if read_bvh_cache and cache.exist_bvh(key, bvh_name):
        logger.info('Loading BVH "%s" for geometry from cache.', bvh_name)
        bvh = cache.load_bvh(key, bvh_name)
    elif auto_build_bvh:
        logger.info('Building new BVH using recursive grid algorithm.')
        start = time.time()
        if cuda_device is not None:
            from .gpu import create_cuda_context
            create_cuda_context(cuda_device)
        bvh = make_recursive_grid_bvh(geometry.mesh, target_degree=3)
        logger.info('BVH generated in %1.1f seconds.', time.time() - start)
        if update_bvh_cache:
            logger.info('Saving BVH (%s:%s) to cache.', key, bvh_name)
            cache.save_bvh(bvh, key, bvh_name)
    return bvh


def create_geometry_from_obj(obj, bvh_name='default', auto_build_bvh=True, read_bvh_cache=True,
                             update_bvh_cache=True, cache_dir=None, cuda_device=None):
    """Geometry (or Detector) with mesh and BVH from a Geometry, Solid, Mesh or a callable
    returning one (chroma/loader.py:162-199)."""
    if callable(obj):
        obj = obj()
    if isinstance(obj, Geometry):                  # Detector is a Geometry
        geometry = obj
    elif isinstance(obj, Solid):
        geometry = Geometry()
        geometry.add_solid(obj)
    elif isinstance(obj, Mesh):
        geometry = Geometry()
        geometry.add_solid(Solid(obj, vacuum, vacuum, color=0x33ffffff))
    else:
        raise TypeError('cannot build type %s' % type(obj))
    geometry.flatten()
    if geometry.bvh is None:
        geometry.bvh = load_bvh(geometry, bvh_name=bvh_name, auto_build_bvh=auto_build_bvh,
                                read_bvh_cache=read_bvh_cache, update_bvh_cache=update_bvh_cache,
                                cache_dir=cache_dir, cuda_device=cuda_device)
    return geometry
