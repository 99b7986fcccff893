"""BVH data model in the reference node format plus the builder entry point.

Node packing (chroma/bvh/bvh.py:8-41, chroma/cuda/geometry_types.h:86-96):
uint4 {x,y,z = lo16 | hi16<<16 ; w = nchild<<28 | child}; leaf <=> nchild == 0 and
child is the triangle id; world = origin + q*scale.  The build itself runs in
libchroma_b200.so (cb_bvh_build: Morton sort on the GPU, layer grouping in C++).
"""
import ctypes as C
import numpy as np

from . import _lib
from .gpuarray import vec

uint4 = vec.uint4
CHILD_BITS = 28
NCHILD_MASK = np.uint32(0xF0000000)


def unpack_nodes(nodes):
    out = np.empty(len(nodes), dtype=[('xlo', np.uint16), ('xhi', np.uint16), ('ylo', np.uint16),
                                      ('yhi', np.uint16), ('zlo', np.uint16), ('zhi', np.uint16),
                                      ('child', np.uint64), ('nchild', np.uint16)])
    for axis in 'xyz':
        out[axis + 'lo'] = nodes[axis] & 0xFFFF
        out[axis + 'hi'] = nodes[axis] >> 16
    out['child'] = nodes['w'] & ~NCHILD_MASK
    out['nchild'] = nodes['w'] >> CHILD_BITS
    return out


class OutOfRangeError(ValueError):
    """World coordinates that do not fit the unsigned 16-bit grid (chroma/bvh/bvh.py:37-42)."""


class WorldCoords(object):
    """world = world_origin + fixed * world_scale (chroma/bvh/bvh.py:44-94)."""

    MAX_INT = 2 ** 16 - 1

    def __init__(self, world_origin, world_scale):
        self.world_origin = np.array(world_origin, dtype=np.float32)
        self.world_scale = np.float32(world_scale)

    def world_to_fixed(self, world):
        fixed = ((np.asarray(world, dtype=np.float64) - self.world_origin) / self.world_scale).round()
        if int(fixed.max()) > self.MAX_INT or fixed.min() < 0:
            raise OutOfRangeError('range = (%f, %f)' % (fixed.min(), fixed.max()))
        return fixed.astype(np.uint16)

    def fixed_to_world(self, fixed):
        return np.asarray(fixed) * self.world_scale + self.world_origin


def node_areas(nodes):
    """Surface area of every node's box in grid units (chroma/bvh/bvh.py:197-212)."""
    u = unpack_nodes(nodes)
    dx, dy, dz = (u[a + 'hi'].astype(np.float64) - u[a + 'lo'] for a in 'xyz')
    return 2.0 * (dx * dy + dy * dz + dz * dx)


class BVHLayerSlice(object):
    """One layer of a BVH as a view of the parent's node array (chroma/bvh/bvh.py:214-260)."""

    def __init__(self, world_coords, nodes):
        self.world_coords = world_coords
        self.nodes = nodes

    def __len__(self):
        return len(self.nodes)

    def areas_fixed(self):
        return node_areas(self.nodes)

    def area_fixed(self):
        return node_areas(self.nodes).sum()

    def area(self):
        return self.area_fixed() * float(self.world_coords.world_scale) ** 2

    def get_bounds(self):
        u = unpack_nodes(self.nodes)
        lower = np.stack([u[a + 'lo'] for a in 'xyz'], axis=-1)
        upper = np.stack([u[a + 'hi'] for a in 'xyz'], axis=-1)
        return (np.atleast_2d(self.world_coords.fixed_to_world(lower)),
                np.atleast_2d(self.world_coords.fixed_to_world(upper)))


class BVH(object):
    """Nodes root first, layers contiguous and in order of depth (chroma/bvh/bvh.py:106-195)."""

    def __init__(self, world_coords, nodes, layer_offsets):
        self.world_coords = world_coords
        self.nodes = nodes
        self.layer_offsets = list(layer_offsets)
        self.layer_bounds = self.layer_offsets + [len(nodes)]

    def get_layer(self, layer_number):
        return BVHLayerSlice(self.world_coords,
                             self.nodes[self.layer_bounds[layer_number]:self.layer_bounds[layer_number + 1]])

    def layer_count(self):
        return len(self.layer_offsets)

    def __len__(self):
        return len(self.nodes)


def make_recursive_grid_bvh(mesh, target_degree=3):
    """Recursive-grid BVH (behaviour of chroma/bvh/grid.py:11-95): one leaf per
    triangle, Morton-ordered, parents formed by dropping low Morton bits until the
    mean fan-out reaches target_degree, <= 15 children, single-child chains collapsed."""
    lib = _lib.lib()
    v = np.ascontiguousarray(mesh.vertices, dtype=np.float32)
    t = np.ascontiguousarray(mesh.triangles, dtype=np.uint32)
    origin = (C.c_float * 3)()
    scale = C.c_float()
    nnodes = C.c_uint64()
    nlayers = C.c_int32()
    _lib.check(lib.cb_bvh_build(v.ctypes.data, len(v), t.ctypes.data, len(t), int(target_degree),
                                C.byref(origin), C.byref(scale), None, C.byref(nnodes), None, C.byref(nlayers)))
    nodes = np.empty(nnodes.value, dtype=uint4)
    layers = np.empty(nlayers.value, dtype=np.uint64)
    _lib.check(lib.cb_bvh_build(v.ctypes.data, len(v), t.ctypes.data, len(t), int(target_degree),
                                C.byref(origin), C.byref(scale), nodes.ctypes.data, C.byref(nnodes),
                                layers.ctypes.data, C.byref(nlayers)))
    return BVH(WorldCoords(np.array(list(origin), dtype=np.float32), scale.value), nodes, layers.astype(np.int64))


def native_tree(nodes, ntriangles, solid_id=None, mesh=None, world_coords=None, max_pieces=0, min_extent=8,
                min_ratio=2.0):
    """The engine's own traversal tree (uint4 entries, root at 0) for a
    reference-format tree; host-only helper around cb_native_tree_build.  With ``mesh``,
    ``world_coords`` and ``max_pieces`` > 1 loosely bounded triangles are referenced by several
    leaves with tighter boxes (cb_native_tree_build_split)."""
    lib = _lib.load()
    nodes = np.ascontiguousarray(nodes)
    sid = None if solid_id is None else np.ascontiguousarray(solid_id, dtype=np.uint32)
    sid_p = sid.ctypes.data if sid is not None else None
    count = C.c_uint64()
    if mesh is not None and max_pieces > 1:
        v = np.ascontiguousarray(mesh.vertices, dtype=np.float32)
        t = np.ascontiguousarray(mesh.triangles, dtype=np.uint32)
        origin = np.ascontiguousarray(world_coords.world_origin, dtype=np.float32)

        def call(out):
            return lib.cb_native_tree_build_split(nodes.ctypes.data, len(nodes), int(ntriangles), sid_p, v.ctypes.data,
                                                  t.ctypes.data, origin.ctypes.data, float(world_coords.world_scale),
                                                  int(max_pieces), int(min_extent), float(min_ratio), out, C.byref(count))
    else:
        def call(out):
            return lib.cb_native_tree_build(nodes.ctypes.data, len(nodes), int(ntriangles), sid_p, out, C.byref(count))
    _lib.check(call(None))
    out = np.empty(count.value, dtype=uint4)
    _lib.check(call(out.ctypes.data))
    return out
