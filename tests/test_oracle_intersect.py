"""CPU oracle for the BVH traversal: pinned to the reference's golden vector and
cross-checked against brute force; reference test order (tie-break rank)."""
import os
import numpy as np

from chroma_lite_b200.geometry import Mesh, Solid, Geometry, vacuum
from chroma_lite_b200.sample import uniform_sphere
from oracle import orc
import scenes

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def brute_force(geo, o, d):
    """float64 Moller-Trumbore over all triangles (independent of the BVH)."""
    tri = geo.mesh.assemble().astype(np.float64)
    v0, e1, e2 = tri[:, 0], tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]
    best_t = np.full(len(o), np.inf)
    best = np.full(len(o), -1)
    for i in range(len(o)):
        dd = d[i].astype(np.float64)
        dd /= np.linalg.norm(dd)
        h = np.cross(dd, e2)
        a = np.einsum('ij,ij->i', e1, h)
        ok = np.abs(a) > 1e-12
        f = np.where(ok, 1.0 / np.where(ok, a, 1.0), 0.0)
        s = o[i].astype(np.float64) - v0
        u = f * np.einsum('ij,ij->i', s, h)
        q = np.cross(s, e1)
        v = f * (q @ dd)
        t = f * np.einsum('ij,ij->i', e2, q)
        ok &= (u >= 0) & (u <= 1) & (v >= 0) & (u + v <= 1) & (t > 1e-6)
        if ok.any():
            t = np.where(ok, t, np.inf)
            best[i] = int(np.argmin(t))
            best_t[i] = t[best[i]]
    return best, best_t


def test_golden_vector_of_the_reference():
    z = np.load(os.path.join(GOLD, 'cube_rays.npz'))
    mesh = Mesh(z['vertices'], z['triangles'], round=False, remove_null_triangles=False)
    geo = Geometry(vacuum)
    geo.add_solid(Solid(mesh, vacuum, vacuum))
    geo.flatten(dedupe_vertices=False)
    scenes.with_bvh(geo)
    desc, keep = scenes.desc_of(geo)
    tri, dist, cnt = orc.intersect(desc, z['pos'], z['dir'])
    gold = z['distance']
    nz = gold != 0
    assert (tri >= 0).all()
    assert np.allclose(dist[nz], gold[nz], rtol=1e-6, atol=0)
    assert cnt['nodes'] == 20 * len(gold)          # 21-node tree, every non-root node visited


def test_oracle_vs_brute_force():
    geo = scenes.sphere_scene(16)
    desc, keep = scenes.desc_of(geo)
    rng = np.random.default_rng(3)
    n = 300
    o = rng.uniform(-3000, 3000, (n, 3)).astype(np.float32)
    d = uniform_sphere(n, rng=rng).astype(np.float32)
    tri, dist, _ = orc.intersect(desc, o, d)
    btri, bt = brute_force(geo, o, d)
    agree = tri == btri
    assert agree.mean() > 0.99
    hit = agree & (tri >= 0)
    assert np.allclose(dist[hit], bt[hit], rtol=1e-4)


def test_rank_is_a_permutation_in_reference_test_order():
    geo = scenes.sphere_scene(16)
    desc, keep = scenes.desc_of(geo)
    rank = orc.triangle_rank(desc)
    assert sorted(rank) == list(range(len(geo.mesh.triangles)))
    # group semantics: the first group's leaf children are tested first, in ascending node order
    nodes = keep['nodes'].view(np.uint32).reshape(-1, 4)
    root = nodes[0, 3]
    first, n = root & 0x0FFFFFFF, root >> 28
    leaf_children = [nodes[i, 3] for i in range(first, first + n) if nodes[i, 3] >> 28 == 0]
    assert [rank[t] for t in leaf_children] == list(range(len(leaf_children)))


def test_last_hit_is_excluded():
    geo = scenes.water_box(100.0)
    desc, keep = scenes.desc_of(geo)
    o = np.zeros((1, 3), np.float32)
    d = np.array([[1.0, 0.2, 0.1]], np.float32)
    t0, d0, _ = orc.intersect(desc, o, d)
    t1, d1, _ = orc.intersect(desc, o, d, last_hit=t0)
    assert t0[0] >= 0 and t1[0] != t0[0]
