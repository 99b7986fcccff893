"""NumPy restatement of the reference's recursive-grid BVH builder.
TEST INFRASTRUCTURE ONLY.

Follows chroma/bvh/grid.py:11-95 (host grouping) with the device kernels
make_leaves / make_parents_detailed / copy_and_offset / collapse_child
(chroma/cuda/bvh.cu:148,269,364,530) restated in oracle/chroma_oracle.c, and
create_leaf_nodes / concatenate_layers / collapse_chains of chroma/gpu/bvh.py.
The Morton argsort here is stable (NumPy's default in the reference is not, so
trees can differ for equal codes)."""
import ctypes as C
import numpy as np

from . import orc

MAX_CHILD = 15
uint4 = np.dtype([('x', np.uint32), ('y', np.uint32), ('z', np.uint32), ('w', np.uint32)])


def count_unique_in_sorted(a):
    return int((np.ediff1d(a) > 0).sum()) + 1


def make_recursive_grid_bvh(vertices, triangles, target_degree=3, leaves=None, parents=None, concatenate=None,
                            collapse=None):
    """Returns (world_origin f32[3], world_scale f32, nodes uint32 (N,4), layer_offsets).
    The four device kernels of the reference's builder can each be plugged in (the GPU test tier passes
    the reference's own kernels from oracle/_ref/bvh.cubin through oracle/ref_driver.py); without a hook
    the C restatement in chroma_oracle.c does the step:
      leaves(v, t, origin, scale) -> (leaf_nodes, morton)        make_leaves            bvh.cu:148
      parents(top, first_child, nchild) -> parent nodes          make_parents_detailed  bvh.cu:269
      concatenate(layers) -> (nodes, bounds)                     copy_and_offset        bvh.cu:364
      collapse(nodes, bounds) -> nodes                           collapse_child         bvh.cu:530"""
    lib = orc.lib()
    v = np.ascontiguousarray(vertices, dtype=np.float32)
    t = np.ascontiguousarray(triangles, dtype=np.uint32)
    world_origin = v.min(axis=0)
    world_scale = np.float32(np.float64(np.max(v.max(axis=0) - world_origin)) / (2 ** 16 - 2))
    leaf = np.zeros((len(t), 4), dtype=np.uint32)
    codes = np.zeros(len(t), dtype=np.uint64)
    if leaves is None:
        lib.orc_make_leaves(v.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p), C.c_uint64(len(t)),
                            world_origin.ctypes.data_as(C.c_void_p), C.c_float(world_scale),
                            leaf.ctypes.data_as(C.c_void_p), codes.ctypes.data_as(C.c_void_p))
    else:
        leaf, codes = leaves(v, t, world_origin, world_scale)
        leaf, codes = np.ascontiguousarray(leaf, dtype=np.uint32), np.ascontiguousarray(codes, dtype=np.uint64)
    order = np.argsort(codes, kind='stable')
    leaf, codes = np.ascontiguousarray(leaf[order]), codes[order]
    layers = [leaf]
    while len(layers[0]) > 1:
        top = layers[0]
        nnodes = len(top)
        nunique = count_unique_in_sorted(codes)
        while nnodes / float(nunique) < target_degree and nunique > 1:
            codes = codes >> np.uint64(1)
            nunique = count_unique_in_sorted(codes)
        delta = np.ediff1d(codes, to_begin=np.uint64(1)).astype(np.uint64)
        parent_codes = codes[delta > 0]
        first_child = np.argwhere(delta > 0).flatten().astype(np.uint32)
        nchild = np.ediff1d(first_child, to_end=nnodes - first_child[-1]).astype(np.uint32)
        if (nchild > MAX_CHILD).any():
            fc, pc = [], []
            for f, n, c in zip(first_child, nchild, parent_codes):
                starts = np.arange(f, f + n, MAX_CHILD, dtype=np.uint32)
                fc.append(starts)
                pc.append(np.repeat(c, len(starts)))
            first_child = np.concatenate(fc)
            parent_codes = np.concatenate(pc).astype(np.uint64)
            nchild = np.ediff1d(first_child, to_end=nnodes - first_child[-1]).astype(np.uint32)
        assert (nchild > 0).all() and (nchild <= MAX_CHILD).all()
        if parents is None:
            layer = np.zeros((len(first_child), 4), dtype=np.uint32)
            lib.orc_make_parents(top.ctypes.data_as(C.c_void_p), first_child.ctypes.data_as(C.c_void_p),
                                 nchild.ctypes.data_as(C.c_void_p), C.c_uint64(len(first_child)),
                                 layer.ctypes.data_as(C.c_void_p))
        else:
            layer = np.ascontiguousarray(parents(top, first_child, nchild), dtype=np.uint32)
        layers = [layer] + layers
        codes = parent_codes
    if concatenate is None:
        bounds = np.insert(np.cumsum([len(l) for l in layers]), 0, 0)
        nodes = np.concatenate(layers).astype(np.uint32)
        for s, e in zip(bounds[:-2], bounds[1:-1]):           # every layer but the leaves
            w = nodes[s:e, 3]
            nodes[s:e, 3] = (w & np.uint32(0xF0000000)) | ((w & np.uint32(0x0FFFFFFF)) + np.uint32(e))
        nodes = np.ascontiguousarray(nodes)
    else:
        nodes, bounds = concatenate(layers)
        nodes = np.ascontiguousarray(nodes, dtype=np.uint32)
    if collapse is None:
        for s, e in reversed(list(zip(bounds[:-2], bounds[1:-1]))):
            lib.orc_collapse_child(nodes.ctypes.data_as(C.c_void_p), C.c_uint64(s), C.c_uint64(e))
    else:
        nodes = np.ascontiguousarray(collapse(nodes, bounds), dtype=np.uint32)
    return world_origin, world_scale, nodes, bounds[:-1]


def attach_bvh(geometry, target_degree=3):
    """Give a flattened geometry a .bvh built by this oracle (CPU only)."""
    from chroma_lite_b200.bvh import BVH, WorldCoords
    if not hasattr(geometry, 'mesh'):
        geometry.flatten()
    o, s, nodes, offs = make_recursive_grid_bvh(geometry.mesh.vertices, geometry.mesh.triangles, target_degree)
    geometry.bvh = BVH(WorldCoords(o, s), nodes.view(uint4)[:, 0], offs)
    return geometry
