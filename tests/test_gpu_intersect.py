"""Ray/mesh nearest hit: engine vs the reference's own intersect_mesh (bit-exact
triangle and distance), vs the CPU oracle, vs the reference's golden vector."""
import os
import numpy as np
import pytest

from chroma_lite_b200 import gpu
from chroma_lite_b200.geometry import Mesh, Solid, Geometry, vacuum
from chroma_lite_b200.sample import uniform_sphere
from oracle import orc, ref_driver
import scenes

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def random_rays(geo, n, seed):
    rng = np.random.default_rng(seed)
    lo, hi = geo.mesh.get_bounds()
    c, h = (lo + hi) / 2, (hi - lo) / 2 * 1.5
    o = (c + rng.uniform(-1, 1, (n, 3)) * h).astype(np.float32)
    d = uniform_sphere(n, rng=rng).astype(np.float32)
    return o, d


def run_engine(geo, o, d, last=None):
    g = gpu.GPUGeometry(geo)
    tri, dist = gpu.intersect_mesh(g, o, d, last)
    return tri.get(), dist.get()


@pytest.mark.parametrize('scene', ['sphere', 'tiny', 'scint'])
def test_bit_exact_vs_reference_kernel(gpu_ready, scene):
    geo = {'sphere': scenes.sphere_scene, 'tiny': scenes.tiny_detector, 'scint': scenes.scintillator_scene}[scene]()
    n = 400000
    o, d = random_rays(geo, n, 1234)
    # also rays from the centre and rays aimed exactly at vertices / edge midpoints (tie-heavy)
    v = geo.mesh.vertices[np.random.default_rng(1).integers(0, len(geo.mesh.vertices), 20000)]
    t = geo.mesh.triangles[np.random.default_rng(2).integers(0, len(geo.mesh.triangles), 20000)]
    mid = 0.5 * (geo.mesh.vertices[t[:, 0]] + geo.mesh.vertices[t[:, 1]])
    src = np.array([3.0, -2.0, 1.0], dtype=np.float32)
    o = np.concatenate([o, np.tile(src, (40000, 1))]).astype(np.float32)
    d = np.concatenate([d, v - src, mid - src]).astype(np.float32)
    tri, dist = run_engine(geo, o, d)
    desc, keep = scenes.desc_of(geo)
    rg = ref_driver.RefGeometry(desc, keep)
    rtri, rdist, _ = ref_driver.intersect(rg, o, d)
    assert np.array_equal(tri, rtri), 'triangle mismatch on %d rays' % (tri != rtri).sum()
    hit = rtri >= 0
    assert np.array_equal(dist[hit].view(np.uint32), rdist[hit].view(np.uint32))
    assert hit.mean() > 0.05


def test_ray_splitting_does_not_change_results(gpu_ready, monkeypatch):
    # the end game of a launch shares the longest rays among the free lanes of their warp
    # (persistent_intersect); with few rays every warp is in that regime from the start
    geo = scenes.tiny_detector()
    for n in (3000, 200000):
        o, d = random_rays(geo, n, 77)
        monkeypatch.setenv('CHROMA_B200_SPLIT', '0')
        tri0, dist0 = run_engine(geo, o, d)
        monkeypatch.setenv('CHROMA_B200_SPLIT', '1')
        tri1, dist1 = run_engine(geo, o, d)
        assert np.array_equal(tri0, tri1)
        hit = tri0 >= 0
        assert np.array_equal(dist0[hit].view(np.uint32), dist1[hit].view(np.uint32))
        assert hit.mean() > 0.05
    monkeypatch.delenv('CHROMA_B200_SPLIT')


def test_last_hit_exclusion_vs_reference(gpu_ready):
    geo = scenes.sphere_scene()
    n = 100000
    o = np.zeros((n, 3), dtype=np.float32)
    d = uniform_sphere(n, rng=np.random.default_rng(5)).astype(np.float32)
    tri0, dist0 = run_engine(geo, o, d)
    # restart from the hit point excluding the triangle just hit
    p = (o + d * dist0[:, None]).astype(np.float32)
    tri1, dist1 = run_engine(geo, p, d, tri0)
    desc, keep = scenes.desc_of(geo)
    rg = ref_driver.RefGeometry(desc, keep)
    rtri, rdist, _ = ref_driver.intersect(rg, p, d, tri0)
    assert np.array_equal(tri1, rtri)
    assert (tri1 != tri0).all()


def test_vs_cpu_oracle(gpu_ready):
    geo = scenes.tiny_detector()
    o, d = random_rays(geo, 50000, 7)
    tri, dist = run_engine(geo, o, d)
    desc, keep = scenes.desc_of(geo)
    otri, odist, cnt = orc.intersect(desc, o, d)
    same = tri == otri
    assert same.mean() > 0.9995          # CPU float arithmetic differs by ulps near ties/edges
    hit = same & (tri >= 0)
    rel = np.abs(dist[hit] - odist[hit]) / odist[hit]
    assert (rel < 1e-5).mean() > 0.999 and rel.max() < 1e-2     # grazing hits are ill-conditioned in float


def test_reference_golden_vector(gpu_ready):
    z = np.load(os.path.join(GOLD, 'cube_rays.npz'))
    mesh = Mesh(z['vertices'], z['triangles'], round=False, remove_null_triangles=False)
    geo = Geometry(vacuum)
    geo.add_solid(Solid(mesh, vacuum, vacuum))
    geo.flatten(dedupe_vertices=False)
    scenes.with_bvh(geo)
    tri, dist = run_engine(geo, z['pos'], z['dir'])
    gold = z['distance']
    nz = gold != 0
    assert (tri >= 0).all()              # incl. the 56 diagonal rays the old test missed
    assert np.allclose(dist[nz], gold[nz], rtol=1e-6)


def test_axis_aligned_and_degenerate_rays(gpu_ready):
    geo = scenes.water_box(100.0)
    o = np.zeros((6, 3), dtype=np.float32)
    d = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], dtype=np.float32)
    tri, dist = run_engine(geo, o, d)
    assert (tri >= 0).all() and np.allclose(dist, 50.0)
    desc, keep = scenes.desc_of(geo)
    rtri, rdist, _ = ref_driver.intersect(ref_driver.RefGeometry(desc, keep), o, d)
    assert np.array_equal(tri, rtri) and np.array_equal(dist, rdist)
    # outside looking away: miss, distance untouched (0)
    tri, dist = run_engine(geo, np.array([[500, 0, 0]], np.float32), np.array([[1, 0, 0]], np.float32))
    assert tri[0] == -1 and dist[0] == 0.0


def test_empty_ray_set(gpu_ready):
    geo = scenes.water_box(10.0)
    tri, dist = run_engine(geo, np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32))
    assert len(tri) == 0 and len(dist) == 0


def test_untraceable_rays_report_no_hit(gpu_ready):
    """A NaN / infinite component or a zero direction makes every slab comparison false: such a ray would
    walk the whole tree (139 ms for ONE photon in the 29k-PMT detector, the 8-GPU straggler of round 2).
    They are not traced and report no hit; the rays around them are unaffected."""
    geo = scenes.ref_tiny_detector()
    g = gpu.GPUDetector(geo)
    ph = scenes.point_source(4096, seed=3)
    o, d = np.array(ph.pos, dtype=np.float32), np.array(ph.dir, dtype=np.float32)
    tri0, dist0 = [a.get() for a in gpu.intersect_mesh(g, o, d)]
    bad = {5: ('d', (np.nan, 0.0, 1.0)), 77: ('d', (0.0, 0.0, 0.0)), 300: ('o', (np.inf, 0.0, 0.0)),
           301: ('o', (0.0, np.nan, 0.0)), 4000: ('d', (np.inf, 1.0, 0.0))}
    for i, (which, v) in bad.items():
        (o if which == 'o' else d)[i] = v
    tri, dist = [a.get() for a in gpu.intersect_mesh(g, o, d)]
    good = np.ones(len(o), dtype=bool)
    good[list(bad)] = False
    assert (tri[list(bad)] == -1).all()
    assert np.array_equal(tri[good], tri0[good]) and np.array_equal(dist[good].view(np.uint32), dist0[good].view(np.uint32))


def test_rays_along_an_axis_do_not_walk_the_whole_slab(gpu_ready, monkeypatch):
    """d == 0 on an axis: the reference skips that axis in its box test (intersect.h:120) and so enters every box in
    the ray's way on the other axes -- for a photon reflected exactly along z (3e-8 of rng_sphere's directions) a slab
    through the whole detector: 164 ms for ONE photon in the 29k-PMT detector, the 8-GPU straggler of round 2.  The
    engine still culls on such an axis (phased_ray_axis); hits are the same (bit-exact test above), the work is not."""
    monkeypatch.setenv('CHROMA_B200_STATS', '1')
    geo = scenes.ref_tiny_detector()
    g = gpu.GPUDetector(geo)
    n = 4096
    rng = np.random.default_rng(5)
    from chroma_lite_b200 import event
    pos = rng.uniform(-1500, 1500, (n, 3)).astype(np.float32)
    work = {}
    for name, d in (('random', None), ('along z', (0.0, 0.0, 1.0)), ('along -x', (-1.0, 0.0, 0.0)), ('in the yz plane', (0.0, 0.6, 0.8))):
        dirs = scenes.point_source(n, seed=8).dir if d is None else np.tile(np.asarray(d, np.float32), (n, 1))
        pol = np.cross(dirs, (0.3, 0.5, 0.81))
        pol /= np.linalg.norm(pol, axis=1)[:, None]
        ph = event.Photons(pos, dirs, pol.astype(np.float32), np.full(n, 400.0, np.float32))
        gp = gpu.GPUPhotons(ph)
        gp.propagate(g, gpu.get_rng_states(n, seed=2), nthreads_per_block=256, max_blocks=16, max_steps=1)
        st = gp.last_stats
        work[name] = st.nodes_visited / float(st.steps)
    print(work)
    for name in ('along z', 'along -x', 'in the yz plane'):
        assert work[name] < 2.0 * work['random'], work
