"""On-disk cache of BVHs (and flattened meshes), role of chroma/cache.py:46-245 with a raw
binary format instead of pickles: a BVH of the 29k-PMT detector is
0.83 GB of nodes; np.load(mmap_mode='r') opens it in milliseconds and an upload reads it once.

Same method names and lookup rule as the reference for the BVH part: BVHs live under
<cache_dir>/bvh/<mesh md5>/<name>, the mesh hash being Mesh.md5() of the reference
(chroma/geometry.py:107-112: md5 of the vertex array, then of the triangle array).
"""
import hashlib
import json
import os
import shutil

import numpy as np

from .bvh import BVH, WorldCoords, uint4

# the reference reads CHROMA_CACHE_DIR (chroma/cache.py:17); its entries are pickles under the same names, so this
# cache lives in its own sub-directory of it
DEFAULT_DIR = (os.path.join(os.environ['CHROMA_CACHE_DIR'], 'b200') if os.environ.get('CHROMA_CACHE_DIR')
               else os.path.join(os.path.expanduser('~'), '.chroma_b200'))


class GeometryNotFoundError(Exception):
    pass


class BVHNotFoundError(Exception):
    pass


def mesh_hash(mesh):
    """MD5 over vertices then triangles, hex (the reference's Mesh.md5())."""
    h = hashlib.md5(np.ascontiguousarray(mesh.vertices))
    h.update(np.ascontiguousarray(mesh.triangles))
    return h.hexdigest()


def _verify_or_create_dir(path):
    if os.path.exists(path) and not os.path.isdir(path):
        raise IOError('Non-directory already exists where a cache directory should go: ' + path)
    os.makedirs(path, exist_ok=True)


class Cache(object):
    def __init__(self, cache_dir=DEFAULT_DIR):
        self.cache_dir = cache_dir
        self.geo_dir = os.path.join(cache_dir, 'geo')
        self.bvh_dir = os.path.join(cache_dir, 'bvh')
        for d in (cache_dir, self.geo_dir, self.bvh_dir):
            _verify_or_create_dir(d)

    # ---- BVHs, keyed by mesh hash and name (chroma/cache.py:180-245)
    def get_bvh_directory(self, mesh_hash):
        return os.path.join(self.bvh_dir, mesh_hash)

    def get_bvh_filename(self, mesh_hash, name='default'):
        return os.path.join(self.get_bvh_directory(mesh_hash), name)

    def list_bvh(self, mesh_hash):
        d = self.get_bvh_directory(mesh_hash)
        return sorted(os.listdir(d)) if os.path.isdir(d) else []

    def exist_bvh(self, mesh_hash, name='default'):
        return os.path.isfile(os.path.join(self.get_bvh_filename(mesh_hash, name), 'meta.json'))

    def save_bvh(self, bvh, mesh_hash, name='default'):
        target = self.get_bvh_filename(mesh_hash, name)
        tmp = target + '.tmp%d' % os.getpid()
        _verify_or_create_dir(tmp)
        nodes = np.ascontiguousarray(bvh.nodes).view(np.uint32).reshape(-1, 4)
        np.save(os.path.join(tmp, 'nodes.npy'), nodes)
        meta = {'format': 1, 'nnodes': int(len(nodes)), 'layer_offsets': [int(x) for x in bvh.layer_offsets],
                'world_origin': [float(x) for x in np.asarray(bvh.world_coords.world_origin)],
                'world_scale': float(bvh.world_coords.world_scale)}
        with open(os.path.join(tmp, 'meta.json'), 'w') as f:
            json.dump(meta, f)
        if os.path.isdir(target):
            shutil.rmtree(target)
        os.replace(tmp, target)                       # readers never see a half-written entry

    def load_bvh(self, mesh_hash, name='default', mmap=True):
        """BVH for the mesh with this hash; `mmap` maps the node array instead of reading it."""
        if not self.exist_bvh(mesh_hash, name):
            raise BVHNotFoundError(mesh_hash + ':' + name)
        d = self.get_bvh_filename(mesh_hash, name)
        with open(os.path.join(d, 'meta.json')) as f:
            meta = json.load(f)
        nodes = np.load(os.path.join(d, 'nodes.npy'), mmap_mode='r' if mmap else None)
        if len(nodes) != meta['nnodes']:
            raise BVHNotFoundError('%s:%s is damaged (%d nodes, header says %d)' % (mesh_hash, name, len(nodes), meta['nnodes']))
        wc = WorldCoords(np.asarray(meta['world_origin'], dtype=np.float32), np.float32(meta['world_scale']))
        return BVH(wc, np.asarray(nodes).view(uint4)[:, 0] if not mmap else nodes.view(uint4)[:, 0], meta['layer_offsets'])

    def remove_bvh(self, mesh_hash, name='default'):
        d = self.get_bvh_filename(mesh_hash, name)
        if os.path.isdir(d):
            shutil.rmtree(d)

    # ---- flattened meshes by name: the arrays GPUGeometry uploads (vertices, triangles, per-triangle
    # material / surface indices, solid ids, colours).  Materials and surfaces are Python objects and
    # stay with the detector description that produced them.
    FLAT_FIELDS = ('vertices', 'triangles', 'colors', 'solid_id', 'material1_index', 'material2_index', 'surface_index')

    def get_geometry_filename(self, name):
        return os.path.join(self.geo_dir, name)

    def list_geometry(self):
        return sorted(n for n in os.listdir(self.geo_dir) if '.tmp' not in n)

    def save_geometry(self, name, geometry):
        """Store the flattened arrays of `geometry` (flatten() is called if needed)."""
        geometry.flatten()
        target = self.get_geometry_filename(name)
        tmp = target + '.tmp%d' % os.getpid()
        _verify_or_create_dir(tmp)
        arrays = {'vertices': geometry.mesh.vertices, 'triangles': geometry.mesh.triangles}
        arrays.update({f: getattr(geometry, f) for f in self.FLAT_FIELDS[2:]})
        for f, a in arrays.items():
            np.save(os.path.join(tmp, f + '.npy'), np.ascontiguousarray(a))
        with open(os.path.join(tmp, 'meta.json'), 'w') as f:
            json.dump({'format': 1, 'mesh_hash': mesh_hash(geometry.mesh), 'ntriangles': int(len(geometry.mesh.triangles))}, f)
        if os.path.isdir(target):
            shutil.rmtree(target)
        os.replace(tmp, target)

    def load_geometry(self, name, mmap=True):
        """dict of the flattened arrays + 'mesh_hash'."""
        d = self.get_geometry_filename(name)
        if not os.path.isfile(os.path.join(d, 'meta.json')):
            raise GeometryNotFoundError(name)
        with open(os.path.join(d, 'meta.json')) as f:
            meta = json.load(f)
        out = {f: np.load(os.path.join(d, f + '.npy'), mmap_mode='r' if mmap else None) for f in self.FLAT_FIELDS}
        out['mesh_hash'] = meta['mesh_hash']
        return out

    def get_geometry_hash(self, name):
        d = self.get_geometry_filename(name)
        if not os.path.isfile(os.path.join(d, 'meta.json')):
            raise GeometryNotFoundError(name)
        with open(os.path.join(d, 'meta.json')) as f:
            return json.load(f)['mesh_hash']

    def remove_geometry(self, name):
        d = self.get_geometry_filename(name)
        if os.path.isdir(d):
            shutil.rmtree(d)

    def load_default_geometry(self, mmap=True):
        """The geometry designated by set_default_geometry (chroma/cache.py:153-160)."""
        return self.load_geometry('.default', mmap=mmap)

    def set_default_geometry(self, name):
        """Point '.default' at the cached geometry ``name`` (a symlink, chroma/cache.py:162-177)."""
        link = self.get_geometry_filename('.default')
        target = self.get_geometry_filename(name)
        if not os.path.isfile(os.path.join(target, 'meta.json')):
            raise GeometryNotFoundError(name)
        if os.path.lexists(link):
            if not os.path.islink(link):
                raise IOError('Non-symlink found where expected a symlink: ' + link)
            os.remove(link)
        os.symlink(target, link)
