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


def test_split_leaves_cover_their_triangles_and_keep_results():
    """Leaf splitting (cb_native_tree_build_split): a triangle may be referenced by several leaves with
    tighter boxes.  Every piece lies inside the reference leaf box, every point of a triangle lies inside
    at least one of its pieces (geometrically nothing is lost), and on these rays the oracle's traversal returns
    the same hits through the split tree.  (Not a proof of equality with the reference: a ray grazing a sliver
    can get a float32-spurious hit only through the full leaf box, see bvh_native.cu and DESIGN section 7-0.)"""
    for geo, pieces, by_solid in ((scenes.tiny_detector(), 8, True), (scenes.tiny_detector(), 8, False),
                                  (scenes.sphere_scene(16), 4, True), (scenes.scintillator_scene(10), 16, True)):
        desc, keep = scenes.desc_of(geo)
        ntri = len(geo.mesh.triangles)
        wc = geo.bvh.world_coords
        sid = geo.solid_id if by_solid else None          # None: one hierarchy over all leaves (CHROMA_B200_TREE=single)
        plain = native_tree(keep['nodes'], ntri, sid)
        nat = native_tree(keep['nodes'], ntri, sid, mesh=geo.mesh, world_coords=wc, max_pieces=pieces,
                          min_extent=4, min_ratio=1.5)
        tris, maxfan, ok = walk(nat)
        tris = np.asarray(tris)
        counts = np.bincount(tris, minlength=ntri)
        assert counts.min() >= 1 and counts.max() <= pieces and maxfan <= 8 and ok
        assert counts.mean() > 1.3, 'nothing was split'
        # pieces inside the reference leaf box
        ref = keep['nodes']
        ref_leaf = unpack_nodes(ref[((ref['w'] >> 28) == 0) & (ref['w'] < ntri)])
        leaf_mask = ((nat['w'] >> 28) == 0) & ((nat['x'] | nat['y'] | nat['z']) != 0)
        pieces_u = unpack_nodes(nat[leaf_mask])
        ptri = pieces_u['child'].astype(np.int64)
        assert set(ptri.tolist()) == set(range(ntri))
        for a in 'xyz':
            lo, hi = np.zeros(ntri, np.int64), np.zeros(ntri, np.int64)
            lo[ref_leaf['child'].astype(np.int64)] = ref_leaf[a + 'lo']
            hi[ref_leaf['child'].astype(np.int64)] = ref_leaf[a + 'hi']
            assert (pieces_u[a + 'lo'] >= lo[ptri]).all() and (pieces_u[a + 'hi'] <= hi[ptri]).all()
        # coverage: random points of every triangle are inside one of its pieces (grid coordinates)
        rng = np.random.default_rng(7)
        verts = geo.mesh.assemble().astype(np.float64)                      # (ntri, 3, 3)
        k = 24
        b = rng.dirichlet(np.ones(3), size=(ntri, k))
        b[:, :3] = np.eye(3)                                                  # the corners themselves
        b[:, 3:6] = (np.eye(3) + np.roll(np.eye(3), 1, axis=1)) / 2          # edge midpoints
        pts = np.einsum('tkc,tcx->tkx', b, verts)
        grid = (pts - wc.world_origin.astype(np.float64)) / float(wc.world_scale)
        # (pieces are clamped to the reference leaf box, whose float32 quantisation can leave a corner a few
        #  1e-4 quanta outside, bvh.cu:65-69; the engine's slab test is widened by ~3 quanta)
        tol = 0.01
        covered = np.zeros((ntri, k), dtype=bool)
        sort = np.argsort(ptri, kind='stable')
        ptri_s, pu = ptri[sort], pieces_u[sort]
        for j in range(counts.max()):                                         # j-th piece of every triangle
            first = np.searchsorted(ptri_s, np.arange(ntri))
            idx = np.minimum(first + j, len(ptri_s) - 1)
            valid = (ptri_s[idx] == np.arange(ntri))
            inside = np.ones((ntri, k), dtype=bool)
            for ax, a in enumerate('xyz'):
                inside &= (grid[:, :, ax] >= pu[a + 'lo'][idx][:, None] - tol) & (grid[:, :, ax] <= pu[a + 'hi'][idx][:, None] + tol)
            covered |= inside & valid[:, None]
        assert covered.all()
        # same nearest hits through the split tree, the unsplit tree and the reference tree
        n = 4000
        lo, hi = geo.mesh.get_bounds()
        o = ((lo + hi) / 2 + rng.uniform(-0.6, 0.6, (n, 3)) * (hi - lo)).astype(np.float32)
        d = uniform_sphere(n, rng=rng).astype(np.float32)
        # plus rays aimed at triangle corners and edge midpoints from the centre of the scene
        aim = pts[rng.integers(0, ntri, 2000), rng.integers(0, 6, 2000)]
        o = np.concatenate([o, np.tile(((lo + hi) / 2 + 1.0).astype(np.float32), (2000, 1))])
        d = np.concatenate([d, (aim - o[n:]).astype(np.float32)])
        t0, d0, _ = orc.intersect(desc, o, d)
        out = []
        for tree in (plain, nat):
            tc = np.ascontiguousarray(tree)
            desc.nodes, desc.nnodes = tc.ctypes.data, len(tc)
            out.append(orc.intersect(desc, o, d))
        assert np.array_equal(d0, out[1][1]) and np.array_equal(out[0][1], out[1][1])
        # (the oracle breaks distance ties by visit order, the engine by reference rank: triangle ids are
        #  compared on the random rays only, the aimed ones tie by construction)
        assert (out[0][0][:n] == out[1][0][:n]).mean() > 0.999 and (t0[:n] == out[1][0][:n]).mean() > 0.999
        assert ((t0 >= 0) == (out[1][0] >= 0)).all()
        assert out[1][2]['tris'] <= out[0][2]['tris']                        # and fewer triangle tests
