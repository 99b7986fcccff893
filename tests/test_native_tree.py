"""The engine's own traversal tree (bvh_native.cu) is built on the host, so its
invariants are checked here without a GPU; the oracle's traversal run on it must
return exactly what it returns on the reference tree."""
import numpy as np

from chroma_lite_b200.bvh import native_tree, unpack_nodes
from chroma_lite_b200.sample import uniform_sphere
from oracle import orc
import scenes


def walk(nodes):
    """(reachable leaf triangles, max fan-out, containment ok)"""
    w = nodes['w']
    u = unpack_nodes(nodes)
    tris, stack, maxfan, ok = [], [0], 0, True
    while stack:
        i = stack.pop()
        n, first = int(w[i] >> 28), int(w[i] & 0x0FFFFFFF)
        if n == 0:
            tris.append(first)
            continue
        maxfan = max(maxfan, n)
        for c in range(first, first + n):
            for a in 'xyz':
                ok &= bool(u[a + 'lo'][c] >= u[a + 'lo'][i] and u[a + 'hi'][c] <= u[a + 'hi'][i])
            stack.append(c)
    return tris, maxfan, ok


def test_native_tree_invariants_and_equivalence():
    for geo in (scenes.sphere_scene(16), scenes.tiny_detector(), scenes.scintillator_scene(10)):
        desc, keep = scenes.desc_of(geo)
        ntri = len(geo.mesh.triangles)
        nat = native_tree(keep['nodes'], ntri, geo.solid_id)
        tris, maxfan, ok = walk(nat)
        assert sorted(tris) == list(range(ntri))          # every triangle exactly once
        assert maxfan <= 8 and ok                          # <= 8 children, parents contain children
        # leaf entries are the reference's leaf boxes verbatim
        ref = keep['nodes']
        ref_leaf = {int(x['w']): (int(x['x']), int(x['y']), int(x['z'])) for x in ref if (x['w'] >> 28) == 0}
        nat_leaf = {int(x['w']): (int(x['x']), int(x['y']), int(x['z'])) for x in nat
                    if (x['w'] >> 28) == 0 and (x['x'] or x['y'] or x['z'])}
        assert all(nat_leaf[t] == ref_leaf[t] for t in range(ntri))
        # same nearest hits through either tree
        rng = np.random.default_rng(4)
        n = 3000
        lo, hi = geo.mesh.get_bounds()
        o = ((lo + hi) / 2 + rng.uniform(-0.6, 0.6, (n, 3)) * (hi - lo)).astype(np.float32)
        d = uniform_sphere(n, rng=rng).astype(np.float32)
        t0, d0, _ = orc.intersect(desc, o, d)
        natc = np.ascontiguousarray(nat)
        desc.nodes, desc.nnodes = natc.ctypes.data, len(natc)
        t1, d1, c1 = orc.intersect(desc, o, d)
        assert np.array_equal(d0, d1) and (t0 == t1).mean() > 0.999   # ties may resolve by visit order
        assert (t0 >= 0).mean() > 0.2


def test_native_tree_without_solids_and_tiny_meshes():
    geo = scenes.water_box(10.0)
    desc, keep = scenes.desc_of(geo)
    nat = native_tree(keep['nodes'], len(geo.mesh.triangles), None)
    tris, maxfan, ok = walk(nat)
    assert sorted(tris) == list(range(12)) and ok
